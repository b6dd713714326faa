// Microbenchmark 4 (round 2): what a narrow-N tcgen05.mma stream costs next to an epilogue that uses shared memory.
//   * cycles per MMA (M = 128, K = 16, SS mode, SWIZZLE_128B K-major operands resident in smem) for N = 16 / 48 / 64 / 144,
//     accumulating into ONE accumulator or alternating between two (is the ~57-cycle floor a dependency latency?)
//   * the same with eight bystander warps streaming LDS.128 + STS.128 (an smem-staged epilogue) or SHFL (a register epilogue):
//     do LSU shared-memory traffic and shuffles share bandwidth with the tensor core's operand reads?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/mma_side tools/exp/mma_side.cu
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
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr int OPER = 64 * 1024;      // operand region: A tiles then B tiles
constexpr int SCRATCH = 64 * 1024;   // bystander region
constexpr int THREADS = 64 + 8 * 32;

// by: 0 none, 1 LDS.128 + STS.128 stream, 2 SHFL stream, 3 LDS.128 only
__global__ void __launch_bounds__(THREADS, 1) side_kernel(int N, int n_mma, int acc_mode, int by, int do_mma, long long* out_cycles, float* sink, long long* by_iters) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* done = (uint64_t*)(smem + OPER + SCRATCH);
  uint32_t* tptr = (uint32_t*)(done + 2);
  volatile int* stop = (volatile int*)(tptr + 2);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (OPER + SCRATCH) / 2; i += blockDim.x)
    ((__nv_bfloat16*)smem)[i] = __float2bfloat16(((i * 2654435761u) >> 24) * (1.f / 256.f) - 0.5f);
  if (threadIdx.x == 0) {
    mbar_init(&done[0], 1);
    *stop = 0;
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
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 32 * 1024);
    long long t0 = clock64();
    if (do_mma) {
      if (elect_one()) {
        for (int i = 0; i < n_mma; i += 8) {
          // walk over two 16 KB A tiles and the B region so that operand fetches are not always the same bytes
          const uint32_t a = a0 + ((i >> 3) & 1) * 16384, b = b0 + ((i >> 3) & 1) * 16384;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const uint32_t acc_off = acc_mode == 1 ? (k & 1) * 256 : 0;
            umma(tmem + acc_off, desc_sw128(a) + 2 * (k & 3), desc_sw128(b) + 2 * (k & 3), idesc, (i | k) != 0 && !(acc_mode == 1 && i == 0 && k == 1));
          }
        }
        commit(&done[0]);
      }
      __syncwarp();
      mbar_wait(&done[0], 0);
    } else {
      while (clock64() - t0 < 400000) {}
    }
    if (lane == 0) { out_cycles[blockIdx.x] = clock64() - t0; *stop = 1; }
  } else if (warp >= 2) {
    const uint32_t base = smem_u32(smem + OPER) + (warp - 2) * 8192 + lane * 16;
    long long iters = 0;
    float acc = 0.f;
    if (by == 1 || by == 3) {
      uint4 v = make_uint4(lane, 1, 2, 3);
      while (!*stop) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 r;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(base + j * 512));
          v.x ^= r.x; v.y += r.y;
          if (by == 1) asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + 4096 + j * 512), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
        }
        iters += 8;
      }
      acc = __uint_as_float(v.x ^ v.y);
    } else if (by == 2) {
      float v = lane;
      while (!*stop) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v += __shfl_down_sync(0xffffffffu, v, 1);
        iters += 16;
      }
      acc = v;
    }
    if (lane == 0) { by_iters[blockIdx.x * 8 + warp - 2] = iters; }
    if (acc == 123.456f) sink[0] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

static void run(int N, int acc_mode, int by, int do_mma) {
  long long *d, *bi; float* sink;
  cudaMalloc(&d, 148 * sizeof(long long)); cudaMalloc(&bi, 148 * 8 * sizeof(long long)); cudaMalloc(&sink, 4);
  cudaMemset(d, 0, 148 * sizeof(long long)); cudaMemset(bi, 0, 148 * 8 * sizeof(long long));
  const int smem = OPER + SCRATCH + 2048;
  cudaFuncSetAttribute(side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int n_mma = 8192;
  side_kernel<<<148, THREADS, smem>>>(N, n_mma, acc_mode, by, do_mma, d, sink, bi);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
  long long h[148], hb[148 * 8];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(hb, bi, sizeof(hb), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
  long long it = 0; for (int w = 0; w < 8; ++w) it += hb[w];
  const double bytes_per_op = by == 1 ? 1024.0 : by == 3 ? 512.0 : by == 2 ? 128.0 : 0.0;
  printf("N=%3d acc=%d bystanders=%s mma=%d: %.1f cycles/MMA (operand bytes/clk %.0f); bystander %.1f B/clk (%.3f warp-ops/clk)\n", N, acc_mode,
         by == 0 ? "none" : by == 1 ? "lds+sts" : by == 2 ? "shfl" : "lds", do_mma, do_mma ? mx / (double)n_mma : 0.0,
         do_mma ? (4096.0 + N * 32.0) / (mx / (double)n_mma) : 0.0, it * bytes_per_op / (double)h[0], it / (double)h[0]);
  cudaFree(d); cudaFree(bi); cudaFree(sink);
}

int main() {
  for (int N : {16, 48, 64, 96, 128, 144, 256})
    for (int acc : {0, 1}) run(N, acc, 0, 1);
  for (int by : {1, 3, 2}) run(48, 0, by, 0);          // bystanders alone
  for (int N : {48, 144})
    for (int by : {1, 3, 2}) run(N, 0, by, 1);
  return 0;
}
