// Microbenchmark 5 (round 2): cycles per tcgen05.mma (kind::f16, bf16 -> fp32, M = 128 per SM, K = 16, SS mode, SWIZZLE_128B
// operands resident in shared memory) as a function of the operand MAJOR-ness -- the weight gradient has the pixel (GEMM-K)
// dimension outermost in memory, so both of its operands are MN-major -- of N, and of cta_group::1 vs cta_group::2
// (M = 256 over the two SMs of a TPC, each SM holding half of B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/mma_major tools/exp/mma_major.cu
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
// K-major SWIZZLE_128B: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t desc_k(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// MN-major SWIZZLE_128B: K rows of 128 B (64 MN elements), 8-row groups 1024 B apart, 64-element MN blocks LBO apart
__device__ __forceinline__ uint64_t desc_mn(uint32_t addr, uint32_t lbo) {
  return (uint64_t)((addr >> 4) & 0x3FFF) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
template <int CG> __device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
template <int CG> __device__ __forceinline__ void commit(uint64_t* bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr int OPER = 192 * 1024;     // A region 64 KB, B region 128 KB
constexpr int THREADS = 64;

// a_mn / b_mn: 1 = MN-major operand.  N = columns of the whole MMA (CG == 2: each CTA holds N / 2 of them).
template <int CG, int a_mn, int b_mn> __global__ void __launch_bounds__(THREADS, 1) major_kernel(int N, int n_acc, int n_mma, long long* out_cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* done = (uint64_t*)(smem + OPER);
  uint32_t* tptr = (uint32_t*)(done + 2);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  for (int i = threadIdx.x; i < OPER / 2; i += blockDim.x)
    ((__nv_bfloat16*)smem)[i] = __float2bfloat16(((i * 2654435761u) >> 24) * (1.f / 256.f) - 0.5f);
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
                           ((uint32_t)((128 * CG) >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 64 * 1024);
    long long t0 = clock64();
    if (rank == 0) {
      if (elect_one()) {
        // descriptors of the two operand regions the stream alternates between, built once: the issue loop only adds the
        // (compile-time) offset of the K step to the start-address field -- a single thread issues every MMA, and a loop that
        // rebuilds 64-bit descriptors with run-time branches costs ~80 cycles per MMA by itself (first version of this file)
        const int n_cta = N / CG;
        uint64_t da0[2], db0[2];
        for (int h = 0; h < 2; ++h) {
          da0[h] = a_mn ? desc_mn(a0 + h * 32768, 16384) : desc_k(a0 + h * 32768);
          db0[h] = b_mn ? desc_mn(b0 + h * 65536, 16384) : desc_k(b0 + h * 65536);
        }
        const uint32_t b_slab = (uint32_t)(n_cta * 128) >> 4;
        uint32_t acc_i = 0;
        for (int i = 0; i < n_mma; i += 8) {
          // MN-major: one 16 KB block = 128 K rows of 64 elements; the 8 steps walk its 16-row slices (2 KB each);
          //   A spans 2 blocks (M = 128), B spans N_cta / 64 blocks, LBO = 16 KB.
          // K-major: one K = 64 slab (rows of 128 B); 4 steps of 32 B inside it, then the next slab (A 16 KB, B N_cta * 128 B further).
          const int h = (i >> 3) & 1;
          const uint32_t d_acc = tmem + acc_i * (uint32_t)N;
          const uint32_t first = i >= 8 * n_acc;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint64_t da = da0[h] + (a_mn ? (uint64_t)(k * 2048 >> 4) : (uint64_t)(((k >> 2) * 16384 + (k & 3) * 32) >> 4));
            const uint64_t db = db0[h] + (b_mn ? (uint64_t)(k * 2048 >> 4) : (uint64_t)((k >> 2) * b_slab + (((k & 3) * 32) >> 4)));
            umma<CG>(d_acc, da, db, idesc, k != 0 ? 1u : first);
          }
          if (++acc_i == (uint32_t)n_acc) acc_i = 0;
        }
        commit<CG>(&done[0]);
      }
      __syncwarp();
    }
    mbar_wait(&done[0], 0);
    if (lane == 0) out_cycles[blockIdx.x] = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  else __syncthreads();
  if (warp == 1) {
    if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
  }
}

template <int CG, int A_MN, int B_MN> static void run(int N, int n_acc) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMemset(d, 0, 148 * sizeof(long long));
  const int smem = OPER + 2048;
  cudaFuncSetAttribute(major_kernel<CG, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 8192;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(148); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CG; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, major_kernel<CG, A_MN, B_MN>, N, n_acc, n_mma, d);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  const double cyc = mx / (double)n_mma;
  printf("cta_group::%d M=%d N=%3d A %s B %s accs=%d: %6.1f cycles/MMA  (tensor work %d cycles; %.0f %% of peak)\n", CG, 128 * CG, N, A_MN ? "MN" : "K ",
         B_MN ? "MN" : "K ", n_acc, cyc, N / 2, 100.0 * (N / 2) / cyc);
  cudaFree(d);
}

template <int CG> static void sweep(int N, int n_acc) {
  run<CG, 0, 0>(N, n_acc); run<CG, 0, 1>(N, n_acc); run<CG, 1, 0>(N, n_acc); run<CG, 1, 1>(N, n_acc);
}

int main() {
  for (int N : {16, 48, 64, 96, 128, 192, 256}) sweep<1>(N, N >= 192 ? 2 : 3);
  run<1, 1, 1>(128, 1);
  for (int N : {64, 128, 192, 256}) sweep<2>(N, N >= 192 ? 2 : 3);
  return 0;
}
