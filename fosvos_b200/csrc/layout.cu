// Frame ingest (NCHW fp32 <-> NHWC activations) and 3x3 weight repacking.
#include "common.cuh"

namespace fosvos {

// One thread per (pixel, 8-channel group): reads are coalesced along W per channel plane,
// writes are 16/32 B vectors.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, T* __restrict__ y, int C, long long HW, int Cp,
                                    long long total /* N*HW*(Cp/8) */) {
  const int groups = Cp / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pix = i % HW;           // pixel fastest -> coalesced plane reads
    long long rest = i / HW;
    int g = (int)(rest % groups);
    long long n = rest / groups;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = g * 8 + j;
      v[j] = c < C ? x[(n * C + c) * HW + pix] : 0.f;
    }
    store8(y + (n * HW + pix) * Cp + g * 8, v);
  }
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ x, float* __restrict__ y, int C, long long HW, int Cp,
                                    long long total) {
  const int groups = Cp / 8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long pix = i % HW;
    long long rest = i / HW;
    int g = (int)(rest % groups);
    long long n = rest / groups;
    float v[8];
    load8(x + (n * HW + pix) * Cp + g * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int c = g * 8 + j;
      if (c < C) y[(n * C + c) * HW + pix] = v[j];
    }
  }
}

// out index -> (co, ci, tap) of the OIHW source, per layout.  The packed tensor is sized by
// the PADDED channel counts (CoutP, CinP: the NHWC strides of the activations); entries whose
// logical channel lies outside the weight are zero.
template <typename T>
__global__ void pack_w_kernel(const float* __restrict__ w, T* __restrict__ out, int Cout, int Cin, int CoutP,
                              int CinP, int layout, int pad, long long total) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    int co, ci, tap;
    if (layout == FOSVOS_W_SIMT_FWD) {            // [tap][ciP][coP]
      co = (int)(i % CoutP);
      ci = (int)((i / CoutP) % CinP);
      tap = (int)(i / ((long long)CoutP * CinP));
    } else if (layout == FOSVOS_W_SIMT_DGRAD) {   // [tap'][coP][ciP], tap = 8 - tap'
      ci = (int)(i % CinP);
      co = (int)((i / CinP) % CoutP);
      tap = 8 - (int)(i / ((long long)CoutP * CinP));
    } else if (layout == FOSVOS_W_TC_FWD) {       // [coP][tap][pad64(ciP)]
      ci = (int)(i % pad);
      tap = (int)((i / pad) % 9);
      co = (int)(i / (9LL * pad));
    } else {                                      // TC_DGRAD: [ciP][tap'][pad64(coP)]
      co = (int)(i % pad);
      tap = 8 - (int)((i / pad) % 9);
      ci = (int)(i / (9LL * pad));
    }
    const bool valid = co < Cout && ci < Cin;
    float v = valid ? w[((long long)co * Cin + ci) * 9 + tap] : 0.f;
    out[i] = from_f32<T>(v);
  }
}

__global__ void pad_bias_kernel(const float* __restrict__ b, float* __restrict__ out, int C, int Cp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Cp) out[i] = (b != nullptr && i < C) ? b[i] : 0.f;
}

static inline int pad64(int c) { return (c + 63) / 64 * 64; }

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_nchw_to_nhwc(const float* x, void* y, int N, int C, int H, int W, int Cp, int dtype,
                        fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nchw_to_nhwc: bad shape");
  FOSVOS_REQUIRE(Cp % 8 == 0 && Cp >= C, "nchw_to_nhwc: Cp=%d must be a multiple of 8 and >= C=%d", Cp, C);
  long long HW = (long long)H * W, total = (long long)N * HW * (Cp / 8);
  int blocks = (int)min((long long)num_sms() * 8, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    nchw_to_nhwc_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(x, (T*)y, C, HW, Cp, total);
  });
  return check_launch("nchw_to_nhwc");
}

int fosvos_nhwc_to_nchw(const void* x, float* y, int N, int C, int H, int W, int Cp, int dtype,
                        fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && y && N > 0 && C > 0 && H > 0 && W > 0, "nhwc_to_nchw: bad shape");
  FOSVOS_REQUIRE(Cp % 8 == 0 && Cp >= C, "nhwc_to_nchw: Cp=%d must be a multiple of 8 and >= C=%d", Cp, C);
  long long HW = (long long)H * W, total = (long long)N * HW * (Cp / 8);
  int blocks = (int)min((long long)num_sms() * 8, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    nhwc_to_nchw_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>((const T*)x, y, C, HW, Cp, total);
  });
  return check_launch("nhwc_to_nchw");
}

long long fosvos_packed_weight_elems(int CoutP, int CinP, int layout) {
  if (CoutP <= 0 || CinP <= 0) return -1;
  switch (layout) {
    case FOSVOS_W_SIMT_FWD:
    case FOSVOS_W_SIMT_DGRAD: return 9LL * CoutP * CinP;
    case FOSVOS_W_TC_FWD: return 9LL * CoutP * pad64(CinP);
    case FOSVOS_W_TC_DGRAD: return 9LL * CinP * pad64(CoutP);
    default: return -1;
  }
}

int fosvos_pack_conv3x3_weight(const float* w, void* out, int Cout, int Cin, int CoutP, int CinP, int layout,
                               int dtype, fosvos_stream_t stream) {
  long long total = fosvos_packed_weight_elems(CoutP, CinP, layout);
  FOSVOS_REQUIRE(w && out && total > 0 && Cout > 0 && Cin > 0 && CoutP >= Cout && CinP >= Cin && CoutP % 8 == 0 &&
                     CinP % 8 == 0,
                 "pack_conv3x3_weight: bad arguments (Cout=%d Cin=%d CoutP=%d CinP=%d layout=%d)", Cout, Cin, CoutP, CinP,
                 layout);
  FOSVOS_REQUIRE(!(layout >= FOSVOS_W_TC_FWD && dtype != FOSVOS_BF16), "pack_conv3x3_weight: TC layouts are bf16 only");
  int pad = layout == FOSVOS_W_TC_FWD ? pad64(CinP) : layout == FOSVOS_W_TC_DGRAD ? pad64(CoutP) : 0;
  int blocks = (int)min((long long)num_sms() * 8, ceil_div_ll(total, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    pack_w_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(w, (T*)out, Cout, Cin, CoutP, CinP, layout, pad, total);
  });
  return check_launch("pack_conv3x3_weight");
}

int fosvos_pad_bias(const float* bias, float* out, int C, int Cp, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(out && C > 0 && Cp >= C, "pad_bias: bad arguments");
  pad_bias_kernel<<<ceil_div(Cp, 256), 256, 0, as_stream(stream)>>>(bias, out, C, Cp);
  return check_launch("pad_bias");
}

}  // extern "C"
