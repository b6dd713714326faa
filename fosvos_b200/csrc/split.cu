// fp32 through the bf16 tensor cores: operand splitting ("bf16x3" with two terms, "bf16x6" with three).
//
// An fp32 value v is carried as bf16 terms x1 = bf16(v), x2 = bf16(v - x1), x3 = bf16(v - x1 - x2): two terms hold 16-17
// significant bits, three hold all 24 (each residual is exact in fp32, so x1 + x2 + x3 == v).  A product of two such
// operands is the sum of the term products; keeping those above 2^-16 (two terms: x1 w1, x2 w1, x1 w2) or 2^-24 (three
// terms: + x3 w1, x2 w2, x1 w3) relative to x1 w1 and accumulating them in fp32 (TMEM) reproduces the fp32 convolution of
// the reference's strict mode (torch fp32 with TF32 off) on tcgen05 -- `conv3x3_tc_kernel<.., SPLIT>` in conv_tc.cu.
//
// Layout: an activation map is ONE bf16 NHWC tensor whose channel dimension holds the terms side by side,
// [x1 | x2 | x3], `seg_len` (a multiple of 64) channels each; the packed weight holds one seg_len-long GEMM-K segment per
// kept product, filled with the weight term of that product (fosvos_split_weight_term).  This file: the kernels that
// produce those layouts (frame ingest, weight packing) and the 2x2 ceil-mode max pool on split maps.
#include "common.cuh"

namespace fosvos {

__device__ __forceinline__ float split_term(float v, int t) {
  for (int i = 0; i < t; ++i) v -= __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// NCHW fp32 frame -> NHWC split map (N, H, W, terms * seg_len); channels >= C of every segment are zero.
// One thread per (pixel, 8-channel group of the OUTPUT): plane reads coalesced along W, 16-byte writes.
__global__ void __launch_bounds__(256)
split_nchw_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int C, long long HW, int seg_len, int terms, long long total) {
  const int groups = terms * seg_len / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i % HW;
    const long long rest = i / HW;
    const int g = (int)(rest % groups);
    const long long n = rest / groups;
    const int t = (g * 8) / seg_len, c0 = g * 8 - t * seg_len;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      v[j] = c < C ? split_term(x[(n * C + c) * HW + pix], t) : 0.f;
    }
    store8(y + (n * HW + pix) * (long long)(terms * seg_len) + g * 8, v);
  }
}

// OIHW fp32 weight -> FOSVOS_W_TC_FWD layout [CoutP][tap][n_pairs * seg_len] whose segment g holds weight term wterm[g]
__global__ void __launch_bounds__(256)
pack_w_split_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int Cout, int Cin, int seg_len, int n_pairs,
                    int wterm_packed, long long total) {
  const int kdim = n_pairs * seg_len;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % kdim);
    const int tap = (int)((i / kdim) % 9);
    const int co = (int)(i / (9LL * kdim));
    const int g = k / seg_len, ci = k - g * seg_len;
    float v = 0.f;
    if (co < Cout && ci < Cin) v = split_term(w[((long long)co * Cin + ci) * 9 + tap], (wterm_packed >> (4 * g)) & 15);
    out[i] = __float2bfloat16_rn(v);
  }
}

// nn.MaxPool2d(2, 2, ceil_mode=True) (osvos_vgg.py:90) on a split map: the terms of a pixel are summed back to the fp32
// value (exact for three terms), the 2x2 window maximum is taken and split again.  One thread per (output pixel, 8 channels).
__global__ void __launch_bounds__(256)
maxpool_split_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int H, int W, int seg_len, int terms,
                     long long total) {
  const int PH = (H + 1) / 2, PW = (W + 1) / 2, groups = seg_len / 8, CT = terms * seg_len;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    long long r = i / groups;
    const int px = (int)(r % PW); r /= PW;
    const int py = (int)(r % PH);
    const long long n = r / PH;
    float best[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) best[j] = -INFINITY;
    for (int dy = 0; dy < 2; ++dy) {
      const int yy = 2 * py + dy;
      if (yy >= H) continue;
      for (int dx = 0; dx < 2; ++dx) {
        const int xx = 2 * px + dx;
        if (xx >= W) continue;
        const __nv_bfloat16* src = x + ((n * H + yy) * (long long)W + xx) * CT + g * 8;
        float v[8], s[8];
        load8(src, s);
        for (int t = 1; t < terms; ++t) {
          load8(src + t * seg_len, v);
#pragma unroll
          for (int j = 0; j < 8; ++j) s[j] += v[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = fmaxf(best[j], s[j]);
      }
    }
    __nv_bfloat16* dst = y + ((n * PH + py) * (long long)PW + px) * CT + g * 8;
    for (int t = 0; t < terms; ++t) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = split_term(best[j], t);
      store8(dst + t * seg_len, v);
    }
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_split_pairs(int terms);
int fosvos_split_weight_term(int terms, int pair);

int fosvos_split_nchw(const float* x_nchw, void* y, int N, int C, int H, int W, int seg_len, int terms, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x_nchw && y && N > 0 && C > 0 && H > 0 && W > 0, "split_nchw: bad shape");
  FOSVOS_REQUIRE(seg_len % 64 == 0 && seg_len >= C && terms >= 1 && terms <= 3, "split_nchw: seg_len=%d (multiple of 64, >= C=%d), terms=%d (1..3)", seg_len, C, terms);
  const long long HW = (long long)H * W, total = (long long)N * HW * (terms * seg_len / 8);
  split_nchw_kernel<<<(int)min((long long)num_sms() * 8, ceil_div_ll(total, 256)), 256, 0, as_stream(stream)>>>(
      x_nchw, (__nv_bfloat16*)y, C, HW, seg_len, terms, total);
  return check_launch("split_nchw");
}

long long fosvos_packed_weight_split_elems(int CoutP, int seg_len, int terms) {
  const int n_pairs = fosvos_split_pairs(terms);
  return (CoutP > 0 && seg_len > 0 && n_pairs > 0) ? 9LL * CoutP * n_pairs * seg_len : -1;
}

int fosvos_pack_conv3x3_weight_split(const float* w_oihw, void* w_packed, int Cout, int Cin, int CoutP, int seg_len, int terms,
                                     fosvos_stream_t stream) {
  const int n_pairs = fosvos_split_pairs(terms);
  FOSVOS_REQUIRE(w_oihw && w_packed && Cout > 0 && Cin > 0 && CoutP >= Cout && CoutP % 8 == 0 && n_pairs > 0 && seg_len % 64 == 0 && seg_len >= Cin,
                 "pack_conv3x3_weight_split: bad arguments (Cout=%d Cin=%d CoutP=%d seg_len=%d terms=%d)", Cout, Cin, CoutP, seg_len, terms);
  int wterm = 0;
  for (int g = 0; g < n_pairs; ++g) wterm |= fosvos_split_weight_term(terms, g) << (4 * g);
  const long long total = fosvos_packed_weight_split_elems(CoutP, seg_len, terms);
  pack_w_split_kernel<<<(int)min((long long)num_sms() * 8, ceil_div_ll(total, 256)), 256, 0, as_stream(stream)>>>(
      w_oihw, (__nv_bfloat16*)w_packed, Cout, Cin, seg_len, n_pairs, wterm, total);
  return check_launch("pack_conv3x3_weight_split");
}

int fosvos_maxpool2x2_split(const void* x, void* y, int N, int H, int W, int seg_len, int terms, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && y && N > 0 && H > 0 && W > 0 && seg_len % 64 == 0 && terms >= 1 && terms <= 3, "maxpool2x2_split: bad arguments");
  const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (seg_len / 8);
  maxpool_split_kernel<<<(int)min((long long)num_sms() * 8, ceil_div_ll(total, 256)), 256, 0, as_stream(stream)>>>(
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, H, W, seg_len, terms, total);
  return check_launch("maxpool2x2_split");
}

}  // extern "C"
