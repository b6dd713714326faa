// 2x2 / stride-2 max pooling with ceil_mode=True over NHWC activations
// (nn.MaxPool2d(kernel_size=2, stride=2, ceil_mode=True), reference osvos_vgg.py:90).
// HBM-bound: one thread per (output pixel, 8-channel vector), 16/32-byte accesses,
// consecutive threads walk consecutive channel vectors -> fully coalesced.
#include <math_constants.h>

#include "common.cuh"

namespace fosvos {

// IDX = unsigned (32-bit index arithmetic: a 64-bit division costs ~100 instructions, three of them per thread were
// more than the pooling itself) whenever the element count allows, long long otherwise.
template <typename T, typename IDX>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int C, int OH, int OW,
                                   long long total) {
  const IDX groups = (IDX)(C / 8);
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < (IDX)total; i += (IDX)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    IDX r = i / groups;
    const int ox = (int)(r % (IDX)OW); r /= (IDX)OW;
    const int oy = (int)(r % (IDX)OH);
    const long long n = (long long)(r / (IDX)OH);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -CUDART_INF_F;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < H && ix < W) {                       // ceil_mode: the last window is clipped
          float v[8];
          load8(x + ((n * H + iy) * W + ix) * C + g * 8, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
        }
      }
    store8(y + ((n * OH + oy) * OW + ox) * C + g * 8, m);
  }
}

// Gradient goes to the FIRST maximum of the window in (row, col) scan order -- the index
// max_pool2d_with_indices records (strict '>' comparison against a running maximum).
template <typename T, typename IDX>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, const T* add, T* dx, int H,
                                   int W, int C, int OH, int OW, long long total) {
  const IDX groups = (IDX)(C / 8);
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < (IDX)total; i += (IDX)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    IDX r = i / groups;
    const int ox = (int)(r % (IDX)OW); r /= (IDX)OW;
    const int oy = (int)(r % (IDX)OH);
    const long long n = (long long)(r / (IDX)OH);
    float v[4][8];
    float m[8];
    int arg[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = -CUDART_INF_F; arg[j] = 0; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
      if (iy < H && ix < W) {
        load8(x + ((n * H + iy) * W + ix) * C + g * 8, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[k][j] > m[j]) { m[j] = v[k][j]; arg[j] = k; }
      }
    }
    float gsrc[8];
    load8(dy + ((n * OH + oy) * OW + ox) * C + g * 8, gsrc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
      if (iy < H && ix < W) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = arg[j] == k ? gsrc[j] : 0.f;
        if (add) {                                    // gradient fan-in: the other consumer's contribution (may alias dx)
          float a[8];
          load8(add + ((n * H + iy) * W + ix) * C + g * 8, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        store8(dx + ((n * H + iy) * W + ix) * C + g * 8, o);
      }
    }
  }
}

// The same with the window index of the maximum already known (two bit planes per 32 channels, written by the fused
// conv + pool epilogues: fosvos_conv3x3_tc_pool_arg): the full-resolution activation is not read at all.
template <typename T, typename IDX>
__global__ void maxpool_bwd_arg_kernel(const uint32_t* __restrict__ arg, const T* __restrict__ dy, const T* add, T* dx, int H,
                                       int W, int C, int OH, int OW, long long total) {
  const IDX groups = (IDX)(C / 8);
  for (IDX i = (IDX)blockIdx.x * blockDim.x + threadIdx.x; i < (IDX)total; i += (IDX)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    IDX r = i / groups;
    const long long opix = (long long)r;               // ((n * OH + oy) * OW + ox)
    const int ox = (int)(r % (IDX)OW); r /= (IDX)OW;
    const int oy = (int)(r % (IDX)OH);
    const long long n = (long long)(r / (IDX)OH);
    const uint2 planes = __ldg(reinterpret_cast<const uint2*>(arg + (opix * (C >> 5) + (g >> 2)) * 2));
    const int sh = (g & 3) * 8;
    const uint32_t lo = planes.x >> sh, hi = planes.y >> sh;
    float gsrc[8];
    load8(dy + opix * C + g * 8, gsrc);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int iy = 2 * oy + (k >> 1), ix = 2 * ox + (k & 1);
      if (iy < H && ix < W) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (int)(((lo >> j) & 1u) | (((hi >> j) & 1u) << 1)) == k ? gsrc[j] : 0.f;
        if (add) {                                    // gradient fan-in: the other consumer's contribution (may alias dx)
          float a[8];
          load8(add + ((n * H + iy) * W + ix) * C + g * 8, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] += a[j];
        }
        store8(dx + ((n * H + iy) * W + ix) * C + g * 8, o);
      }
    }
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "maxpool2x2_fwd: bad shape (C=%d must be a multiple of 8)", C);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long total = (long long)N * OH * OW * (C / 8);
  const int blocks = (int)min((long long)num_sms() * 16, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    if (total + (long long)blocks * 256 < (1LL << 32))
      maxpool_fwd_kernel<T, unsigned><<<blocks, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, H, W, C, OH, OW, total);
    else
      maxpool_fwd_kernel<T, long long><<<blocks, 256, 0, as_stream(stream)>>>((const T*)x, (T*)y, H, W, C, OH, OW, total);
  });
  return check_launch("maxpool2x2_fwd");
}

int fosvos_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C, int dtype,
                          fosvos_stream_t stream) {
  return fosvos_maxpool2x2_bwd_add(x, dy, nullptr, dx, N, H, W, C, dtype, stream);
}

int fosvos_maxpool2x2_bwd_add(const void* x, const void* dy, const void* add, void* dx, int N, int H, int W, int C, int dtype,
                              fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0, "maxpool2x2_bwd: bad shape (C=%d must be a multiple of 8)", C);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long total = (long long)N * OH * OW * (C / 8);
  const int blocks = (int)min((long long)num_sms() * 16, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    if (total + (long long)blocks * 256 < (1LL << 32))
      maxpool_bwd_kernel<T, unsigned><<<blocks, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)dy, (const T*)add, (T*)dx, H, W, C, OH, OW, total);
    else
      maxpool_bwd_kernel<T, long long><<<blocks, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)dy, (const T*)add, (T*)dx, H, W, C, OH, OW, total);
  });
  return check_launch("maxpool2x2_bwd");
}

int fosvos_maxpool2x2_bwd_arg(const void* pool_arg, const void* dy, const void* add, void* dx, int N, int H, int W, int C, int dtype,
                              fosvos_stream_t stream) {
  FOSVOS_REQUIRE(pool_arg && dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && C % 32 == 0 && ((uintptr_t)pool_arg & 7) == 0,
                 "maxpool2x2_bwd_arg: bad arguments (C=%d must be a multiple of 32, the index map 8-byte aligned)", C);
  const int OH = (H + 1) / 2, OW = (W + 1) / 2;
  const long long total = (long long)N * OH * OW * (C / 8);
  const int blocks = (int)min((long long)num_sms() * 16, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    if (total + (long long)blocks * 256 < (1LL << 32))
      maxpool_bwd_arg_kernel<T, unsigned><<<blocks, 256, 0, as_stream(stream)>>>((const uint32_t*)pool_arg, (const T*)dy, (const T*)add, (T*)dx, H, W, C, OH, OW, total);
    else
      maxpool_bwd_arg_kernel<T, long long><<<blocks, 256, 0, as_stream(stream)>>>((const uint32_t*)pool_arg, (const T*)dy, (const T*)add, (T*)dx, H, W, C, OH, OW, total);
  });
  return check_launch("maxpool2x2_bwd_arg");
}

}  // extern "C"
