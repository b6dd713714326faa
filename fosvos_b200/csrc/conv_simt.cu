// Direct (CUDA-core, fp32-accumulate) 3x3 convolution kernels over NHWC activations.
//
// These are the exact-arithmetic companions of the tcgen05 kernels in conv_tc.cu: they
// serve the fp32 precision mode, arbitrary channel counts, and the weight gradient.
// Implicit GEMM with a shared-memory halo patch; every thread owns a 4 pixel x 8 channel
// register tile.
#include "common.cuh"

namespace fosvos {

constexpr int TH = 8, TW = 16;        // output pixel patch per CTA
constexpr int PH = TH + 2, PW = TW + 2;
constexpr int PWP = PW + 1;           // padded row pitch of the halo patch
constexpr int BN = 64;                // output channels per CTA
constexpr int KC = 8;                 // input channels per K step

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_simt_kernel(const T* __restrict__ x, const T* __restrict__ w, const float* __restrict__ bias,
                    const T* __restrict__ mask, T* __restrict__ y, int N, int H, int W, int Cin, int Cout,
                    int flags, int tiles_x, int tiles_y) {
  __shared__ float patch[KC][PH][PWP];
  __shared__ __align__(16) float wsm[9][KC][BN];

  const int tid = threadIdx.x;
  int t = blockIdx.x;
  const int tx = t % tiles_x; t /= tiles_x;
  const int ty = t % tiles_y; t /= tiles_y;
  const int n = t;
  const int y0 = ty * TH, x0 = tx * TW;
  const int co0 = blockIdx.y * BN;

  const int cg = tid & 7;          // 8-channel group inside the 64-channel tile
  const int pg = tid >> 3;         // 0..31: pixel group (row, 4-pixel run)
  const int row = pg >> 2, xs = (pg & 3) * 4;

  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[j][c] = 0.f;

  const T* xn = x + (long long)n * H * W * Cin;

  for (int ci0 = 0; ci0 < Cin; ci0 += KC) {
    // ---- stage the halo patch: one 8-channel vector per halo pixel
    for (int i = tid; i < PH * PW; i += 256) {
      const int py = i / PW, px = i % PW;
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float v[8];
      if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
        load8(xn + ((long long)gy * W + gx) * Cin + ci0, v);   // Cin % 8 == 0
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) patch[j][py][px] = v[j];
    }
    // ---- stage the weights: [tap][ci][co0..co0+64)
    for (int i = tid; i < 9 * KC * (BN / 8); i += 256) {
      const int v8 = i % (BN / 8);
      const int ci = (i / (BN / 8)) % KC;
      const int tap = i / (BN / 8 * KC);
      const int co = co0 + v8 * 8;
      float v[8];
      if (co < Cout) {
        load8(w + ((long long)tap * Cin + ci0 + ci) * Cout + co, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      *reinterpret_cast<float4*>(&wsm[tap][ci][v8 * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&wsm[tap][ci][v8 * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();

#pragma unroll
    for (int ci = 0; ci < KC; ++ci) {
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float p[6];
#pragma unroll
        for (int j = 0; j < 6; ++j) p[j] = patch[ci][row + r][xs + j];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float4 wa = *reinterpret_cast<const float4*>(&wsm[r * 3 + s][ci][cg * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&wsm[r * 3 + s][ci][cg * 8 + 4]);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[j][c] = fmaf(p[j + s], wv[c], acc[j][c]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue
  const int co = co0 + cg * 8;
  if (co >= Cout) return;
  float b[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) b[c] = (flags & FOSVOS_CONV_BIAS) ? bias[co + c] : 0.f;
  const int gy = y0 + row;
  if (gy >= H) return;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int gx = x0 + xs + j;
    if (gx >= W) continue;
    const long long o = (((long long)n * H + gy) * W + gx) * Cout + co;
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = acc[j][c] + b[c];
      if (flags & FOSVOS_CONV_RELU) v[c] = fmaxf(v[c], 0.f);
    }
    if (flags & FOSVOS_CONV_MASK) {
      float m[8];
      load8(mask + o, m);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = m[c] > 0.f ? v[c] : 0.f;
    }
    if (flags & FOSVOS_CONV_ACCUMULATE) {
      float old[8];
      load8(y + o, old);
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] += old[c];
    }
    store8(y + o, v);
  }
}

// ---- weight gradient ---------------------------------------------------------------
// dw[co][ci][tap] += sum_p x[p+tap][ci] * dz[p][co]   (GEMM with K = pixels)
// CTA tile: 32 input channels x 64 output channels x 9 taps; K walks 8x8 pixel patches,
// split across blockIdx.z; fp32 atomics merge the splits into the parameter's .grad.
constexpr int WT = 8;                 // wgrad pixel patch edge
constexpr int WPH = WT + 2, WPW = WT + 3;
constexpr int WCI = 32, WCO = 64;

template <typename T>
__global__ void __launch_bounds__(256)
conv3x3_wgrad_simt_kernel(const T* __restrict__ x, const T* __restrict__ dz, float* __restrict__ dw,
                          float* __restrict__ db, int N, int H, int W, int Cin, int Cout, int CinW, int CoutW,
                          int tiles_x, int tiles_y) {
  __shared__ float xs[WCI][WPH][WPW];
  __shared__ __align__(16) float dzs[WT * WT][WCO];

  const int tid = threadIdx.x;
  const int ci0 = blockIdx.x * WCI, co0 = blockIdx.y * WCO;
  const int cg = tid & 7, ci = tid >> 3;

  float acc[9][8];
#pragma unroll
  for (int a = 0; a < 9; ++a)
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[a][c] = 0.f;
  float bsum[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) bsum[c] = 0.f;
  const bool do_bias = (db != nullptr) && blockIdx.x == 0 && ci == 0;

  const int n_tiles = N * tiles_x * tiles_y;
  for (int tile = blockIdx.z; tile < n_tiles; tile += gridDim.z) {
    int t = tile;
    const int tx = t % tiles_x; t /= tiles_x;
    const int ty = t % tiles_y; t /= tiles_y;
    const int n = t;
    const int y0 = ty * WT, x0 = tx * WT;
    const T* xn = x + (long long)n * H * W * Cin;
    const T* dn = dz + (long long)n * H * W * Cout;

    for (int i = tid; i < WPH * (WT + 2) * (WCI / 8); i += 256) {
      const int v8 = i % (WCI / 8);
      const int pp = i / (WCI / 8);
      const int py = pp / (WT + 2), px = pp % (WT + 2);
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      const int c = ci0 + v8 * 8;
      float v[8];
      if (gy >= 0 && gy < H && gx >= 0 && gx < W && c < Cin) {
        load8(xn + ((long long)gy * W + gx) * Cin + c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) xs[v8 * 8 + j][py][px] = v[j];
    }
    for (int i = tid; i < WT * WT * (WCO / 8); i += 256) {
      const int v8 = i % (WCO / 8);
      const int pp = i / (WCO / 8);
      const int gy = y0 + pp / WT, gx = x0 + pp % WT;
      const int c = co0 + v8 * 8;
      float v[8];
      if (gy < H && gx < W && c < Cout) {
        load8(dn + ((long long)gy * W + gx) * Cout + c, v);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
      }
      *reinterpret_cast<float4*>(&dzs[pp][v8 * 8]) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(&dzs[pp][v8 * 8 + 4]) = make_float4(v[4], v[5], v[6], v[7]);
    }
    __syncthreads();

#pragma unroll 2
    for (int pp = 0; pp < WT * WT; ++pp) {
      const int py = pp / WT, px = pp % WT;
      const float4 da = *reinterpret_cast<const float4*>(&dzs[pp][cg * 8]);
      const float4 dbv = *reinterpret_cast<const float4*>(&dzs[pp][cg * 8 + 4]);
      const float dv[8] = {da.x, da.y, da.z, da.w, dbv.x, dbv.y, dbv.z, dbv.w};
      if (do_bias) {
#pragma unroll
        for (int c = 0; c < 8; ++c) bsum[c] += dv[c];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float xv = xs[ci][py + r][px + s];
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r * 3 + s][c] = fmaf(xv, dv[c], acc[r * 3 + s][c]);
        }
    }
    __syncthreads();
  }

  const int cin = ci0 + ci;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int co = co0 + cg * 8 + c;
    if (co >= CoutW) continue;
    if (cin < CinW) {
#pragma unroll
      for (int a = 0; a < 9; ++a) atomicAdd(dw + ((long long)co * CinW + cin) * 9 + a, acc[a][c]);
    }
    if (do_bias) atomicAdd(db + co, bsum[c]);
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_conv3x3_simt(const void* x, const void* w, const float* bias, const void* mask, void* y, int N,
                        int H, int W, int Cin, int Cout, int flags, int dtype, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && w && y && N > 0 && H > 0 && W > 0, "conv3x3_simt: null pointer or empty shape");
  FOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cin > 0 && Cout > 0,
                 "conv3x3_simt: Cin=%d and Cout=%d must be positive multiples of 8 (pad the NHWC tensors)", Cin, Cout);
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_BIAS) || bias, "conv3x3_simt: BIAS flag without bias pointer");
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_MASK) || mask, "conv3x3_simt: MASK flag without mask pointer");
  const int tiles_x = ceil_div(W, TW), tiles_y = ceil_div(H, TH);
  dim3 grid(tiles_x * tiles_y * N, ceil_div(Cout, BN));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    conv3x3_simt_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)w, bias, (const T*)mask, (T*)y,
                                                                N, H, W, Cin, Cout, flags, tiles_x, tiles_y);
  });
  return check_launch("conv3x3_simt");
}

int fosvos_conv3x3_wgrad_simt(const void* x, const void* dz, float* dw, float* db, int N, int H, int W, int Cin,
                              int Cout, int CinW, int CoutW, int dtype, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && dz && dw && N > 0 && H > 0 && W > 0, "conv3x3_wgrad_simt: null pointer or empty shape");
  FOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cin > 0 && Cout > 0,
                 "conv3x3_wgrad_simt: CinP=%d and CoutP=%d must be positive multiples of 8", Cin, Cout);
  FOSVOS_REQUIRE(CinW > 0 && CinW <= Cin && CoutW > 0 && CoutW <= Cout, "conv3x3_wgrad_simt: logical dims exceed padded dims");
  const int tiles_x = ceil_div(W, WT), tiles_y = ceil_div(H, WT);
  const int n_tiles = N * tiles_x * tiles_y;
  const int gx = ceil_div(Cin, WCI), gy = ceil_div(Cout, WCO);
  int split = max(1, min(n_tiles, (num_sms() * 4) / (gx * gy)));
  dim3 grid(gx, gy, split);
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    conv3x3_wgrad_simt_kernel<T><<<grid, 256, 0, as_stream(stream)>>>((const T*)x, (const T*)dz, dw, db, N, H, W, Cin,
                                                                      Cout, CinW, CoutW, tiles_x, tiles_y);
  });
  return check_launch("conv3x3_wgrad_simt");
}

}  // extern "C"
