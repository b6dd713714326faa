// Shared helpers for the fosvos_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fosvos_b200.h"

namespace fosvos {

// ---- error plumbing (no exceptions across the ABI) --------------------------------
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define FOSVOS_REQUIRE(cond, ...)        \
  do {                                   \
    if (!(cond)) {                       \
      ::fosvos::set_error(__VA_ARGS__);  \
      return FOSVOS_ERR_BAD_ARG;         \
    }                                    \
  } while (0)

static inline cudaStream_t as_stream(fosvos_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
int num_sms();
// Per-device one-time setup (cudaFuncSetAttribute is per device): true exactly when `done` had no bit for the CURRENT
// device yet; sets it.  A benign race at worst repeats an idempotent attribute call.
int device_slot();
static inline bool first_use_on_device(unsigned long long& done) {
  const unsigned long long bit = 1ull << device_slot();
  if (done & bit) return false;
  done |= bit;
  return true;
}

// ---- element conversion -------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// 8 consecutive elements <-> 8 floats (16 B for bf16, 32 B for fp32); p must be 16B aligned.
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  uint4 r = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 r;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fosvos

// dispatch on the activation dtype
#define FOSVOS_DISPATCH_DTYPE(dtype, T, ...)                          \
  do {                                                                \
    if ((dtype) == FOSVOS_F32) {                                      \
      using T = float;                                                \
      __VA_ARGS__                                                     \
    } else if ((dtype) == FOSVOS_BF16) {                              \
      using T = __nv_bfloat16;                                        \
      __VA_ARGS__                                                     \
    } else {                                                          \
      ::fosvos::set_error("bad dtype %d", (int)(dtype));              \
      return FOSVOS_ERR_BAD_ARG;                                      \
    }                                                                 \
  } while (0)
