// Microbenchmark 2: the conv kernel's barrier protocol without any data movement: a producer thread that recycles
// smem stages through empty/full mbarriers, the MMA issuer, and (optionally) spinning bystander warps.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try(bar, parity)) {} }
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) { while (!mbar_try(bar, parity)) { __nanosleep(200); } }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int STAGE_MAX = 48 * 1024;
constexpr int MAXST = 4;

// mode bit0: producer/consumer barrier ring; bit1: bystander warps spin (try_wait); bit2: bystanders use nanosleep backoff;
// bit3: fence after each full wait
__global__ void __launch_bounds__(192, 1) pipe_kernel(int N, int n_kb, int stages, int mode, int STAGE, int acc_mode, long long* out_cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + MAXST * STAGE_MAX);
  uint64_t* empty = full + MAXST;
  uint64_t* done = empty + MAXST;
  uint32_t* tptr = (uint32_t*)(done + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < MAXST * STAGE_MAX / 2; i += blockDim.x)
    ((__nv_bfloat16*)smem)[i] = __float2bfloat16(((i * 2654435761u) >> 24) * (1.f / 256.f) - 0.5f);
  if (threadIdx.x == 0) {
    for (int i = 0; i < MAXST; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(&done[0], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tptr)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tptr;
  if (warp == 0) {
    if (lane == 0 && (mode & 1)) {
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        mbar_arrive(&full[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      long long t0 = clock64();
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < n_kb; ++kb) {
        if (mode & 1) {
          mbar_wait(&full[stage], phase);
          if (mode & 8) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const uint32_t a = smem_u32(smem + stage * STAGE);
        const uint64_t da = desc_sw128(a), db = desc_sw128(a + 16384);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma(tmem + (acc_mode == 1 ? (k & 1) * 256 : acc_mode == 2 ? (kb & 1) * 256 : acc_mode == 3 ? ((kb >> 2) & 1) * 256 : 0), da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
        commit(&empty[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      commit(&done[0]);
      mbar_wait(&done[0], 0);
      out_cycles[blockIdx.x] = clock64() - t0;
    }
  } else {
    if (mode & 2) {
      if (mode & 4) mbar_wait_sleep(&done[0], 0); else mbar_wait(&done[0], 0);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static void run(int N, int n_kb, int stages, int mode, int STAGE, int acc_mode) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMemset(d, 0, 148 * sizeof(long long));
  const int smem = MAXST * STAGE_MAX + 2048;
  cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int it = 0; it < 3; ++it) {
    cudaEventRecord(e0);
    pipe_kernel<<<148, 192, smem>>>(N, n_kb, stages, mode, STAGE, acc_mode, d);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e2)); exit(1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("N=%3d stage=%dK acc=%d stages=%d mode=%2d (ring=%d spin=%d sleep=%d fence=%d): %.1f cycles/MMA, kernel %.1f us\n", N, STAGE / 1024, acc_mode, stages, mode, mode & 1,
         (mode >> 1) & 1, (mode >> 2) & 1, (mode >> 3) & 1, mx / (4.0 * n_kb), best * 1e3);
  cudaFree(d);
}

int main() {
  const int n_kb = 2048;
  for (int N : {64, 128, 256})
    for (int stage : {24, 32, 48})
      for (int acc : {0, 1, 2, 3}) {
        if (stage * 1024 < 16384 + N * 128) continue;
        run(N, n_kb, 4, 0, stage * 1024, acc);
      }
  for (int N : {64, 128, 256})
    for (int mode : {1, 9, 3, 7, 11}) run(N, n_kb, 4, mode, 48 * 1024, 0);
  for (int st : {2, 3}) run(64, n_kb, st, 11, 48 * 1024, 0);
  return 0;
}
