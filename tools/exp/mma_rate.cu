// Microbenchmark: issue rate of tcgen05.mma (kind::f16, bf16, M=128 per CTA) as a function of N, of the number of
// MMAs between commits, and of cta_group (1 or 2).  Operands are SWIZZLE_128B K-major tiles in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/mma_rate tools/exp/mma_rate.cu
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int CG>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)1) : "memory");
}

// stage = A tile (128 x 64 bf16 = 16 KB) + B tile (256 x 64 bf16 = 32 KB)
constexpr int STAGE = 48 * 1024;
constexpr int NSTAGE = 4;

template <int CG>
__global__ void __launch_bounds__(128, 1) rate_kernel(int N, int n_kb, int per_commit, long long* out_cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bar = (uint64_t*)(smem + NSTAGE * STAGE);
  uint32_t* tptr = (uint32_t*)(bar + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  // fill operands with small bf16 values
  for (int i = threadIdx.x; i < NSTAGE * STAGE / 2; i += blockDim.x)
    ((__nv_bfloat16*)smem)[i] = __float2bfloat16(((i * 2654435761u) >> 24) * (1.f / 256.f) - 0.5f);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && lane == 0 && rank == 0) {
    const int Mi = (CG == 2) ? 256 : 128;
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(Mi >> 4) << 24);
    t0 = clock64();
    uint32_t phase = 0;
    int since = 0;
    for (int kb = 0; kb < n_kb; ++kb) {
      const uint32_t a = smem_u32(smem + (kb % NSTAGE) * STAGE);
      const uint64_t da = desc_sw128(a), db = desc_sw128(a + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma<CG>(tmem + ((kb >> 2) & 1) * 256, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
      if (++since == per_commit) {
        since = 0;
        commit<CG>(&bar[0]);
        if (per_commit >= 1000) {}  // never
      }
    }
    commit<CG>(&bar[1]);
    mbar_wait(&bar[1], 0);
    t1 = clock64();
    (void)phase;
    out_cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG == 2) { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
  if (warp == 0) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG>
static void run(int N, int n_kb, int per_commit, int grid) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMemset(d, 0, 148 * sizeof(long long));
  const int smem = NSTAGE * STAGE + 2048;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 4; ++it) {
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, n_kb, per_commit, d);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("CG=%d N=%d error %s / %s\n", CG, N, cudaGetErrorString(e), cudaGetErrorString(e2)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  const double mmas = 4.0 * n_kb;
  const double flops = 2.0 * (CG == 2 ? 256 : 128) * N * 16 * mmas * (CG == 2 ? grid / 2 : grid);
  printf("CG=%d N=%3d per_commit=%4d grid=%3d: %.1f cycles/MMA (max CTA), kernel %.1f us, %.0f TFLOP/s\n", CG, N, per_commit, grid,
         mx / mmas, best * 1e3, flops / (best * 1e-3) / 1e12);
  cudaFree(d);
}

int main() {
  const int n_kb = 2048;
  for (int N : {16, 32, 64, 128, 256}) run<1>(N, n_kb, 1 << 30, 148);
  for (int N : {64, 128, 256}) run<1>(N, n_kb, 1, 148);
  for (int N : {64, 256}) run<1>(N, n_kb, 1 << 30, 1);
  for (int N : {32, 64, 128, 256}) run<2>(N, n_kb, 1 << 30, 148);
  for (int N : {64, 128, 256}) run<2>(N, n_kb, 1, 148);
  return 0;
}
