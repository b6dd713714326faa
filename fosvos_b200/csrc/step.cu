// Optimizer-step companions of the tensor-core path, one launch each for ALL conv layers:
//   wgrad_fold_kernel : .grad (OIHW fp32) += live [tap][M][N] weight-gradient accumulators; accumulators = 0
//   repack_kernel     : refresh the packed bf16 weight copies (forward [co][tap][ci64] and data-gradient
//                       [ci][8-tap][co64] layouts) and the padded fp32 bias from the updated fp32 parameters
// (reference: optimizer.step() / zero_grad() at train_online.py:99-100; the packed copies are derived data).
// Both are HBM-bound tile transposes through shared memory: every global access is a contiguous run.
#include <stdlib.h>

#include "common.cuh"

namespace fosvos {

constexpr int FOLD_CO = 32, FOLD_CI = 32;           // tile: 32 couts x 32 cins x 9 taps (36 KB in flight per block)
constexpr int PACK_CO = 32, PACK_CI = 32;           // tile: 32 couts x 32 cins x 9 taps

// tile t of the flat list belongs to entry e with prefix[e] <= t < prefix[e + 1]
__device__ __forceinline__ int find_entry(const int* __restrict__ prefix, int n, int t) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (__ldg(prefix + mid) <= t) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256)
wgrad_fold_kernel(const fosvos_fold_entry* __restrict__ table, int n_entries, const int* __restrict__ prefix, int n_tiles) {
  // row pitch FOLD_CI * 9 + 1 (odd): the gather below writes columns of this array when the workspace is [tap][cin][cout]
  constexpr int FPITCH = FOLD_CI * 9 + 1;
  __shared__ float sm[FOLD_CO * FPITCH];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ei = find_entry(prefix, n_entries, tile);
    const fosvos_fold_entry e = table[ei];
    const int local = tile - __ldg(prefix + ei);
    const int ci_tiles = (e.Cin + FOLD_CI - 1) / FOLD_CI;
    const int co0 = (local / ci_tiles) * FOLD_CO, ci0 = (local % ci_tiles) * FOLD_CI;
    const int nco = min(FOLD_CO, e.Cout - co0), nci = min(FOLD_CI, e.Cin - ci0);
    const long long plane = (long long)e.CoutP * e.CinP;
    // gather: fastest index follows the workspace's contiguous dimension
    for (int i = threadIdx.x; i < 9 * FOLD_CO * FOLD_CI; i += 256) {
      int tap, co_l, ci_l;
      if (e.x_is_a) { co_l = i % FOLD_CO; ci_l = (i / FOLD_CO) % FOLD_CI; tap = i / (FOLD_CO * FOLD_CI); }
      else          { ci_l = i % FOLD_CI; co_l = (i / FOLD_CI) % FOLD_CO; tap = i / (FOLD_CO * FOLD_CI); }
      if (co_l < nco && ci_l < nci) {
        const int co = co0 + co_l, ci = ci0 + ci_l;
        float* src = e.ws + tap * plane + (e.x_is_a ? ((long long)ci * e.CoutP + co) : ((long long)co * e.CinP + ci));
        sm[co_l * FPITCH + ci_l * 9 + tap] = *src;
        *src = 0.f;
      }
    }
    __syncthreads();
    // scatter: per cout a contiguous run of nci * 9 floats of the OIHW gradient
    const int run = nci * 9;
    for (int i = threadIdx.x; i < nco * run; i += 256) {
      const int co_l = i / run, r = i - co_l * run;
      e.dw[((long long)(co0 + co_l) * e.Cin + ci0) * 9 + r] += sm[co_l * FPITCH + r];
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
repack_kernel(const fosvos_repack_entry* __restrict__ table, int n_entries, const int* __restrict__ prefix, int n_tiles) {
  constexpr int PITCH = PACK_CI * 9 + 1;            // odd pitch: conflict-free column reads
  __shared__ float sm[PACK_CO * PITCH];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ei = find_entry(prefix, n_entries, tile);
    const fosvos_repack_entry e = table[ei];
    const int local = tile - __ldg(prefix + ei);
    const int ci_tiles = (e.Cin + PACK_CI - 1) / PACK_CI;
    const int co0 = (local / ci_tiles) * PACK_CO, ci0 = (local % ci_tiles) * PACK_CI;
    const int nco = min(PACK_CO, e.Cout - co0), nci = min(PACK_CI, e.Cin - ci0);
    const int run = nci * 9;
    for (int i = threadIdx.x; i < nco * run; i += 256) {
      const int co_l = i / run, r = i - co_l * run;
      sm[co_l * PITCH + r] = e.w[((long long)(co0 + co_l) * e.Cin + ci0) * 9 + r];
    }
    if (ci0 == 0 && e.bias_out && threadIdx.x < nco)
      e.bias_out[co0 + threadIdx.x] = e.bias ? e.bias[co0 + threadIdx.x] : 0.f;
    __syncthreads();
    __nv_bfloat16* fwd = reinterpret_cast<__nv_bfloat16*>(e.out_fwd);
    __nv_bfloat16* dgr = reinterpret_cast<__nv_bfloat16*>(e.out_dgrad);
    if (fwd) {
      // [co][tap][pad_ci]: ci fastest; a thread writes eight consecutive cins as one 16-byte store (pad_ci and ci0 are
      // multiples of 32, so the address is aligned); channels past the weight's Cin keep their zero padding
      for (int i = threadIdx.x; i < nco * 9 * (PACK_CI / 8); i += 256) {
        const int c8 = i % (PACK_CI / 8), tap = (i / (PACK_CI / 8)) % 9, co_l = i / ((PACK_CI / 8) * 9);
        __nv_bfloat16* dst = fwd + ((long long)(co0 + co_l) * 9 + tap) * e.pad_ci + ci0 + 8 * c8;
        if (8 * c8 + 8 <= nci) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sm[co_l * PITCH + (8 * c8 + q) * 9 + tap];
          store8(dst, v);
        } else {
          for (int q = 0; 8 * c8 + q < nci; ++q) dst[q] = __float2bfloat16_rn(sm[co_l * PITCH + (8 * c8 + q) * 9 + tap]);
        }
      }
    }
    if (dgr) {
      // [ci][8 - tap][pad_co]: co fastest, eight consecutive couts per 16-byte store
      for (int i = threadIdx.x; i < nci * 9 * (PACK_CO / 8); i += 256) {
        const int c8 = i % (PACK_CO / 8), tap = (i / (PACK_CO / 8)) % 9, ci_l = i / ((PACK_CO / 8) * 9);
        __nv_bfloat16* dst = dgr + ((long long)(ci0 + ci_l) * 9 + (8 - tap)) * e.pad_co + co0 + 8 * c8;
        if (8 * c8 + 8 <= nco) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sm[(8 * c8 + q) * PITCH + ci_l * 9 + tap];
          store8(dst, v);
        } else {
          for (int q = 0; 8 * c8 + q < nco; ++q) dst[q] = __float2bfloat16_rn(sm[(8 * c8 + q) * PITCH + ci_l * 9 + tap]);
        }
      }
    }
    __syncthreads();
  }
}

// The whole optimizer step of a 3x3 conv weight in ONE pass over its data (fold + SGD + repack fused): per 32 x 32 x 9 tile
//   g = ws[tap][M][N] (+ .grad, when the caller keeps one);  ws = 0 (.grad = 0)
//   buf = mu * buf + (g + wd * p);  p -= lr * buf                      (torch.optim.SGD, train_online.py:99-100)
//   packed bf16 forward / data-gradient copies and the padded bias refreshed from the new p
// Moves 28 B per parameter (ws read + zeroed, p and buf read + written, 2 x 2 B packed) instead of the 48 B of the three
// separate launches.
__global__ void __launch_bounds__(256, 3)
conv_step_kernel(const fosvos_convstep_entry* __restrict__ table, int n_entries, const int* __restrict__ prefix, int n_tiles, float mu,
                 int allow_vec) {
  constexpr int PITCH = PACK_CI * 9 + 1;
  __shared__ float sm[PACK_CO * PITCH];
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int ei = find_entry(prefix, n_entries, tile);
    const fosvos_convstep_entry e = table[ei];
    const int local = tile - __ldg(prefix + ei);
    const int ci_tiles = (e.Cin + PACK_CI - 1) / PACK_CI;
    const int co0 = (local / ci_tiles) * PACK_CO, ci0 = (local % ci_tiles) * PACK_CI;
    const int nco = min(PACK_CO, e.Cout - co0), nci = min(PACK_CI, e.Cin - ci0);
    const long long plane = (long long)e.CoutP * e.CinP;
    // Full tiles (all but the channel tails and the first layer) move 16 bytes per access: the workspace rows, the OIHW runs
    // (32 x 9 floats per cout) and hence every global access of steps 1 and 2 are float4-aligned there.
    const bool vec = allow_vec && nco == PACK_CO && nci == PACK_CI && (e.Cin & 3) == 0 && (e.CinP & 3) == 0 && (e.CoutP & 3) == 0 &&
                     (((uintptr_t)e.ws | (uintptr_t)e.w | (uintptr_t)e.buf | (uintptr_t)e.dw) & 15) == 0;
    if (vec) {
      // 1. nine float4 per thread (one per tap), all loads before the stores that clear them
      static_assert(PACK_CO == 32 && PACK_CI == 32, "vector path: 8 float4 per row, 32 rows = 256 threads");
      const int q = threadIdx.x & 7, row = threadIdx.x >> 3;          // row = cout (x_is_a = 0) or cin (x_is_a = 1) of the tile
      float* base = e.ws + (e.x_is_a ? ((long long)(ci0 + row) * e.CoutP + co0) : ((long long)(co0 + row) * e.CinP + ci0)) + 4 * q;
      float4 g4[9];
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) g4[tap] = *reinterpret_cast<const float4*>(base + tap * plane);
      // step 2's operands (32 runs of 288 contiguous floats = 2304 float4, nine per thread in three batches) do not depend on
      // step 1: the first batch is requested now, every further one while its predecessor is processed, so a tile costs
      // ~3 dependent DRAM round trips instead of 5 (the kernel is latency-bound: 3 blocks of 8 warps per SM)
      float4 pv[3], bv[3];
      long long gi[3];
      int si[3];
      auto request = [&](int b, float4 (&p4)[3], float4 (&b4)[3], long long (&g)[3], int (&sidx)[3]) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int idx = threadIdx.x + 256 * (3 * b + u);
          const int co_l = idx / 72, r4 = idx - co_l * 72;
          g[u] = ((long long)(co0 + co_l) * e.Cin + ci0) * 9 + 4 * r4;
          sidx[u] = co_l * PITCH + 4 * r4;
          p4[u] = *reinterpret_cast<const float4*>(e.w + g[u]);
          b4[u] = *reinterpret_cast<const float4*>(e.buf + g[u]);
        }
      };
      request(0, pv, bv, gi, si);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) {
        *reinterpret_cast<float4*>(base + tap * plane) = make_float4(0.f, 0.f, 0.f, 0.f);
        const float gv[4] = {g4[tap].x, g4[tap].y, g4[tap].z, g4[tap].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int co_l = e.x_is_a ? 4 * q + j : row, ci_l = e.x_is_a ? row : 4 * q + j;
          sm[co_l * PITCH + ci_l * 9 + tap] = gv[j];
        }
      }
      __syncthreads();
      // 2. SGD with momentum; the new weights replace the gradient in smem
      const float lr = e.lr, wd = e.weight_decay;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        float4 pn4[3], bn4[3];
        long long gn[3];
        int sn[3];
        if (b < 2) request(b + 1, pn4, bn4, gn, sn);
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const float4 dv = e.dw ? *reinterpret_cast<const float4*>(e.dw + gi[u]) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float p4[4] = {pv[u].x, pv[u].y, pv[u].z, pv[u].w}, b4[4] = {bv[u].x, bv[u].y, bv[u].z, bv[u].w};
          const float d4[4] = {dv.x, dv.y, dv.z, dv.w};
          float bn[4], pn[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float gv = sm[si[u] + j] + d4[j];
            bn[j] = mu * b4[j] + (gv + wd * p4[j]);
            pn[j] = p4[j] - lr * bn[j];
            sm[si[u] + j] = pn[j];
          }
          *reinterpret_cast<float4*>(e.buf + gi[u]) = make_float4(bn[0], bn[1], bn[2], bn[3]);
          *reinterpret_cast<float4*>(e.w + gi[u]) = make_float4(pn[0], pn[1], pn[2], pn[3]);
          if (e.dw) *reinterpret_cast<float4*>(e.dw + gi[u]) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (b < 2) {
#pragma unroll
          for (int u = 0; u < 3; ++u) { pv[u] = pn4[u]; bv[u] = bn4[u]; gi[u] = gn[u]; si[u] = sn[u]; }
        }
      }
    } else {
    // 1. gather the accumulator tile (fastest index = the workspace's contiguous dimension) and clear it.  Loads are issued
    // in batches of six BEFORE the stores that clear them: the compiler cannot move a load across an earlier store to the
    // same array, and one load in flight per thread left the kernel latency-bound (2.4 TB/s)
    constexpr int GB = 6;
    static_assert((9 * PACK_CO * PACK_CI) % (256 * GB) == 0, "gather batches");
    for (int i0 = threadIdx.x; i0 < 9 * PACK_CO * PACK_CI; i0 += 256 * GB) {
      float v[GB];
      float* src[GB];
      int dst[GB];
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        const int i = i0 + 256 * u;
        int tap, co_l, ci_l;
        if (e.x_is_a) { co_l = i % PACK_CO; ci_l = (i / PACK_CO) % PACK_CI; tap = i / (PACK_CO * PACK_CI); }
        else          { ci_l = i % PACK_CI; co_l = (i / PACK_CI) % PACK_CO; tap = i / (PACK_CO * PACK_CI); }
        const bool ok = co_l < nco && ci_l < nci;
        const int co = co0 + co_l, ci = ci0 + ci_l;
        src[u] = ok ? e.ws + tap * plane + (e.x_is_a ? ((long long)ci * e.CoutP + co) : ((long long)co * e.CinP + ci)) : nullptr;
        dst[u] = co_l * PITCH + ci_l * 9 + tap;
        v[u] = ok ? *src[u] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        if (src[u]) {
          sm[dst[u]] = v[u];
          *src[u] = 0.f;
        }
      }
    }
    __syncthreads();
    // 2. SGD with momentum over the OIHW runs (nci * 9 contiguous floats per cout); the new weights replace the gradient in
    // smem.  Same batching: four elements' loads before their stores
    const int run = nci * 9, total = nco * run;
    const float lr = e.lr, wd = e.weight_decay;
    constexpr int SB = 4;
    for (int i0 = threadIdx.x; i0 < total; i0 += 256 * SB) {
      float pv[SB], bv[SB], dv[SB];
      long long gi[SB];
      int si[SB];
#pragma unroll
      for (int u = 0; u < SB; ++u) {
        const int i = i0 + 256 * u;
        const bool ok = i < total;
        const int co_l = ok ? i / run : 0, r = ok ? i - co_l * run : 0;
        gi[u] = ok ? ((long long)(co0 + co_l) * e.Cin + ci0) * 9 + r : -1;
        si[u] = co_l * PITCH + r;
        pv[u] = ok ? e.w[gi[u]] : 0.f;
        bv[u] = ok ? e.buf[gi[u]] : 0.f;
        dv[u] = (ok && e.dw) ? e.dw[gi[u]] : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SB; ++u) {
        if (gi[u] >= 0) {
          const float gv = sm[si[u]] + dv[u];
          const float b = mu * bv[u] + (gv + wd * pv[u]);
          const float pn = pv[u] - lr * b;
          e.buf[gi[u]] = b;
          e.w[gi[u]] = pn;
          if (e.dw) e.dw[gi[u]] = 0.f;
          sm[si[u]] = pn;
        }
      }
    }
    }
    if (ci0 == 0 && e.bias_out && threadIdx.x < nco)
      e.bias_out[co0 + threadIdx.x] = e.bias ? e.bias[co0 + threadIdx.x] : 0.f;
    __syncthreads();
    // 3. packed copies (as repack_kernel)
    __nv_bfloat16* fwd = reinterpret_cast<__nv_bfloat16*>(e.out_fwd);
    __nv_bfloat16* dgr = reinterpret_cast<__nv_bfloat16*>(e.out_dgrad);
    if (fwd) {
      for (int i = threadIdx.x; i < nco * 9 * (PACK_CI / 8); i += 256) {
        const int c8 = i % (PACK_CI / 8), tap = (i / (PACK_CI / 8)) % 9, co_l = i / ((PACK_CI / 8) * 9);
        __nv_bfloat16* dst = fwd + ((long long)(co0 + co_l) * 9 + tap) * e.pad_ci + ci0 + 8 * c8;
        if (8 * c8 + 8 <= nci) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sm[co_l * PITCH + (8 * c8 + q) * 9 + tap];
          store8(dst, v);
        } else {
          for (int q = 0; 8 * c8 + q < nci; ++q) dst[q] = __float2bfloat16_rn(sm[co_l * PITCH + (8 * c8 + q) * 9 + tap]);
        }
      }
    }
    if (dgr) {
      for (int i = threadIdx.x; i < nci * 9 * (PACK_CO / 8); i += 256) {
        const int c8 = i % (PACK_CO / 8), tap = (i / (PACK_CO / 8)) % 9, ci_l = i / ((PACK_CO / 8) * 9);
        __nv_bfloat16* dst = dgr + ((long long)(ci0 + ci_l) * 9 + (8 - tap)) * e.pad_co + co0 + 8 * c8;
        if (8 * c8 + 8 <= nco) {
          float v[8];
#pragma unroll
          for (int q = 0; q < 8; ++q) v[q] = sm[(8 * c8 + q) * PITCH + ci_l * 9 + tap];
          store8(dst, v);
        } else {
          for (int q = 0; 8 * c8 + q < nco; ++q) dst[q] = __float2bfloat16_rn(sm[(8 * c8 + q) * PITCH + ci_l * 9 + tap]);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_conv_step_all(const fosvos_convstep_entry* table, int n_entries, const int* tile_prefix, int n_tiles, float momentum,
                         fosvos_stream_t stream) {
  FOSVOS_REQUIRE(table && tile_prefix && n_entries > 0 && n_tiles > 0, "conv_step_all: bad arguments");
  static const int allow_vec = getenv("FOSVOS_STEP_NO_VEC") ? 0 : 1;          // A/B switch (timing experiments)
  conv_step_kernel<<<min(n_tiles, num_sms() * 6), 256, 0, as_stream(stream)>>>(table, n_entries, tile_prefix, n_tiles, momentum, allow_vec);
  return check_launch("conv_step_all");
}

int fosvos_fold_tile_count(int Cout, int Cin) { return ceil_div(Cout, FOLD_CO) * ceil_div(Cin, FOLD_CI); }
int fosvos_repack_tile_count(int Cout, int Cin) { return ceil_div(Cout, PACK_CO) * ceil_div(Cin, PACK_CI); }

int fosvos_wgrad_fold_all(const fosvos_fold_entry* table, int n_entries, const int* tile_prefix, int n_tiles,
                          fosvos_stream_t stream) {
  FOSVOS_REQUIRE(table && tile_prefix && n_entries > 0 && n_tiles > 0, "wgrad_fold_all: bad arguments");
  wgrad_fold_kernel<<<min(n_tiles, num_sms() * 8), 256, 0, as_stream(stream)>>>(table, n_entries, tile_prefix, n_tiles);
  return check_launch("wgrad_fold_all");
}

int fosvos_repack_all(const fosvos_repack_entry* table, int n_entries, const int* tile_prefix, int n_tiles,
                      fosvos_stream_t stream) {
  FOSVOS_REQUIRE(table && tile_prefix && n_entries > 0 && n_tiles > 0, "repack_all: bad arguments");
  repack_kernel<<<min(n_tiles, num_sms() * 6), 256, 0, as_stream(stream)>>>(table, n_entries, tile_prefix, n_tiles);
  return check_launch("repack_all");
}

}  // extern "C"
