// 3x3 convolution with 64 output channels (conv1_2 forward, the data gradients of conv1_2 and conv2_1) on the tensor
// cores, with the three taps of a kernel ROW stacked along GEMM-N -- conv_side_tc.cu's row stack at Cout = 64.
//
// Why: one M = 128 tcgen05.mma costs max(N / 2, 32 + N / 4) cycles (tools/exp/mma_side.cu, mma_major.cu: reading the
// 128-row A tile from shared memory alone takes 32 cycles), so the generic kernel's N = 64 tiles run the tensor pipe
// at 32 / 48 = 67 % at best (43 % measured on conv1_2).  Here
//
//   D[q, (s, co)] = sum_r sum_c X[q + (r - 1, 0), c] * W[co, r, s, c]           one GEMM, N = 3 * 64 = 192, K = 3 * Cin
//   out[y, x, co] = b[co] + D[(y, x - 1), (0, co)] + D[(y, x), (1, co)] + D[(y, x + 1), (2, co)]
//
// an N = 192 instruction is tensor-bound (96 cycles for 96 of work) and replaces three N = 64 ones (144 cycles).
// The three partial sums of an output pixel sit in neighbouring accumulator rows = neighbouring LANES of one warp and
// are combined with two shuffles per channel; results go from registers straight to global memory (32-byte stores, made
// contiguous over lane pairs by a register exchange), so the epilogue uses no shared memory, whose
// traffic would compete with the tensor core's operand reads (80 of the 96 cycles of an N = 192 MMA).
//
// Tile = 16 x 8 pixel patch whose first and last columns are halo (14 x 8 outputs, 87.5 % useful rows): GEMM row
// m = 16 * row + column, so a warp (32 TMEM lanes) holds two image rows -- lanes l and l ^ 16 are vertical
// neighbours, and the fused 2x2 ceil-mode max pool of conv1_2 (osvos_vgg.py:90) is one shuffle down the column and
// one along the row (tile origins are even in both directions).  A operand: ONE halo box {64 ch, 16, 8 + 2} per
// 64-channel slab (TMA, SWIZZLE_128B; out-of-frame pixels zero-filled = the conv padding); vertical tap r is the
// same tile read 16 * r rows further (2048 B, a multiple of the swizzle atom).  B operand: ALL weights of the layer
// ([slab][r][(s, co)][64 ch]: 72 KB for Cin = 64, 144 KB for Cin = 128) are loaded once per persistent CTA and stay
// resident.  Two TMEM accumulators (256 columns apart); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-9 / 10-17 =
// two epilogue groups draining alternate tiles (two warps per TMEM lane quadrant, 32 channels each).
// Epilogues: forward = bias, ReLU, bf16 store and / or pooled store; data gradient = ReLU mask (mask[n, y, x, co] > 0).
//
// What bounds it (ncu, conv1_2 pool-only at batch 16: 405 us against 526 us for the generic kernel, tensor pipe 72 %): the
// MIO data pipe (`l1tex__data_pipe_lsu_wavefronts` 80 %).  A tile's 128 x 192 fp32 accumulator is 768 wavefronts of
// TMEM -> register traffic (three times the generic kernel's) and the shuffles add as many, together more than the
// 1152 tensor cycles of a one-slab tile.  Two-slab tiles (conv2_1's data gradient: 72 us against 93 us) hide it.  Layers
// that also read a mask at one slab (conv1_2's data gradient) stay on the generic kernel (conv_tc.cu dispatch).
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fosvos {

constexpr int SK_THREADS = 64 + 16 * 32;
constexpr int SK_TW = 16, SK_TH = 8;
constexpr int SK_OUT_W = SK_TW - 2;                       // output columns per tile
constexpr int SK_A_BYTES = (SK_TH + 2) * SK_TW * 128;     // 20480: halo box of one 64-channel slab
constexpr int SK_WT_BYTES = 192 * 128;                    // 24576: weights of one (slab, kernel row): 192 rows (s, co) x 64 ch
constexpr int SK_ACC_STRIDE = 256;                        // TMEM columns between the two accumulators
constexpr int SK_TMEM_COLS = 512;
constexpr int SK_MISC_BYTES = 1024;                       // bias[64], barriers, tmem pointer

struct SkParams {
  const float* bias;             // 64 fp32 or null
  const __nv_bfloat16* mask;     // (N, H, W, 64) or null
  __nv_bfloat16* y;              // (N, H, W, 64) or null (pool-only)
  __nv_bfloat16* y_pool;         // (N, ceil(H/2), ceil(W/2), 64) or null
  uint32_t* pool_arg;            // (N, ceil(H/2), ceil(W/2), 2, 2) or null: window index of the first maximum, two bit planes per 32 channels
  int relu;
  int N, H, W;
  int tiles_x, tiles_y, total_tiles;
  int k_chunks;                  // 64-channel slabs
  int cin_pad;                   // k_chunks * 64: per-tap K extent of the packed weight
  int stages;                    // depth of the halo-box ring
  uint32_t dx_mul, dx_shift, dy_mul, dy_shift;
};

__global__ void __launch_bounds__(SK_THREADS, 1)
conv3x3_stack_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const SkParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                              // stages x SK_A_BYTES
  uint8_t* wres = smem + p.stages * SK_A_BYTES;                      // k_chunks x 3 x SK_WT_BYTES
  uint8_t* misc = wres + p.k_chunks * 3 * SK_WT_BYTES;
  float* bias_s = reinterpret_cast<float*>(misc);                    // 64
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(misc + 256);      // up to 8 stages
  uint64_t* empty_bar = full_bar + 8;
  uint64_t* tmem_full = empty_bar + 8;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* w_bar = tmem_empty + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_w);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 8);            // one arrive per warp of the epilogue group
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, SK_TMEM_COLS);
  if (threadIdx.x < 64) bias_s[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // the layer's weights, once: tile (slab kb, kernel row r) = rows (s, co) from the packed [cout][tap][cin_pad] layout,
    // one {64 ch, 64 couts} box per tap
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_bar, (uint32_t)(p.k_chunks * 3 * SK_WT_BYTES));
      for (int kb = 0; kb < p.k_chunks; ++kb)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s)
            ptx::tma_load_2d(wres + ((kb * 3 + r) * 3 + s) * 8192, &map_w, w_bar, (3 * r + s) * p.cin_pad + kb * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m2 = (int)ptx::fast_div((uint32_t)tile, p.dx_mul, p.dx_shift);
      const int tx = tile - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int x0 = tx * SK_OUT_W - 1, y0 = ty * SK_TH;
      for (int kb = 0; kb < p.k_chunks; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[stage], SK_A_BYTES);
          ptx::tma_load_4d(ring + stage * SK_A_BYTES, &map_x, &full_bar[stage], kb * 64, x0, y0 - 1, n);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 192);
    const uint64_t desc0 = ptx::umma_desc_sw128_kmajor(ptx::smem_u32(ring));
    const uint32_t a_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
    const uint32_t w_lo0 = (uint32_t)ptx::umma_desc_sw128_kmajor(ptx::smem_u32(wres));
    int stage = 0, it = 0;
    uint32_t phase = 0;
    ptx::mbar_wait(w_bar, 0);                           // weights are resident
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * SK_ACC_STRIDE;
      for (int kb = 0; kb < p.k_chunks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_lo = a_lo0 + stage * (SK_A_BYTES >> 4);
          const uint32_t w_lo = w_lo0 + kb * (3 * SK_WT_BYTES >> 4);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(tmem_d, a_lo + r * ((SK_TW * 128) >> 4) + 2 * k, w_lo + r * (SK_WT_BYTES >> 4) + 2 * k, desc_hi, idesc,
                                  (uint32_t)(kb | r | k));
          }
          ptx::umma_commit(&empty_bar[stage]);          // frees the halo box when the MMAs retire
          if (kb == p.k_chunks - 1) ptx::umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: group g (warps 2 + 4 g ..) drains accumulator g = every other tile =====================
    // sixteen warps: group (accumulator = every other tile) x channel half (32 of the 64 output channels) x TMEM lane quadrant.
    // The epilogue is a long dependent chain per thread (TMEM load -> shuffles -> convert -> stores); with one warp per
    // quadrant and group it took 2.3x the MMA time of a one-slab tile, so each quadrant's work is split over two warps.
    const int e = warp - 2;
    const int grp = e >> 3;
    const int cbase = 2 * ((e >> 2) & 1);               // first of this warp's two 16-channel chunks
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access = image rows 2 quad, 2 quad + 1
    const int as = grp;
    const int col = lane & 15, half = lane >> 4;
    const int PH = (p.H + 1) >> 1, PW = (p.W + 1) >> 1;
    int it = grp;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, it += 2) {
      const int m2 = (int)ptx::fast_div((uint32_t)tile, p.dx_mul, p.dx_shift);
      const int tx = tile - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int gx = tx * SK_OUT_W - 1 + col, gy = ty * SK_TH + 2 * quad + half;
      const bool valid = col >= 1 && col <= SK_OUT_W && gx < p.W && gy < p.H;
      const long long pix = ((long long)n * p.H + gy) * p.W + gx;
      // pooled output: written by the top-left pixel of each 2x2 window (even row = lower half-warp, even x = odd column)
      const bool pool_writer = p.y_pool && valid && half == 0 && (col & 1);
      const long long ppix = ((long long)n * PH + (gy >> 1)) * PW + (gx >> 1);
      // Global accesses are made contiguous inside lane PAIRS (two horizontally adjacent pixels, 2 x 64 B of this warp's
      // channel half): access k of lane 2g + j touches chunk cbase + j of pixel 2g + k, so the two lanes cover 64 contiguous
      // bytes and a warp instruction 16 half-lines -- lane-per-pixel accesses (32 lines of 32 B per instruction) kept the LSU
      // as busy as the tensor core.  Registers are exchanged with one shuffle per register.
      const int j2 = lane & 1, g2 = lane & ~1;
      const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
      // the ReLU mask (chunk cbase + j2 of the two pixels of the pair), requested BEFORE waiting for the accumulator: its DRAM
      // latency hides behind the tile's MMAs
      uint32_t mk[2][8];
      if (p.mask) {
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          if ((vmask >> (g2 + k)) & 1u) {
            ptx::ldg256_nc(p.mask + (pix - j2 + k) * 64 + (cbase + j2) * 16, mk[k]);
          } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) mk[k][r] = 0u;
          }
        }
      }
      ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t mbits[2] = {0u, 0u};                      // bit e of mbits[c]: mask[own pixel][16 (cbase + c) + e] > 0
      if (p.mask) {
        // 16 "positive" bits per (pixel k, chunk cbase + j2), both pixels in one word, then handed to the lanes that own the pixels
        uint32_t w = 0u;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          uint32_t bits = 0u;
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            const uint32_t mw = mk[k][r];                  // bf16 > 0: sign bit clear and not (+)zero
            bits |= (((mw & 0x8000u) == 0 && (mw & 0x7fffu) != 0) ? 1u : 0u) << (2 * r);
            bits |= (((mw & 0x80000000u) == 0 && (mw & 0x7fff0000u) != 0) ? 1u : 0u) << (2 * r + 1);
          }
          w |= bits << (16 * k);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) mbits[c] = (__shfl_sync(0xffffffffu, w, g2 + c) >> (16 * j2)) & 0xffffu;
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * SK_ACC_STRIDE;
      uint32_t o[2][8];
      uint32_t arg_lo = 0u, arg_hi = 0u;                  // bit planes of the pool's window index, this warp's 32 channels
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int c0 = 16 * (cbase + c);
        uint32_t d0[16], d1[16], d2[16];
        ptx::tmem_ld16(taddr + c0, d0);
        ptx::tmem_ld16(taddr + 64 + c0, d1);
        ptx::tmem_ld16(taddr + 128 + c0, d2);
        ptx::tmem_ld_wait();
        if (c == 1) {                                         // last read of the accumulator: hand it back before the arithmetic
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[j]), 1);       // kernel column 0: from pixel x - 1
          const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[j]), 1);    // kernel column 2: from pixel x + 1
          v[j] = (left + __uint_as_float(d1[j])) + (right + bias_s[c0 + j]);
        }
        if (p.mask) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            o[c][j] = ptx::cvt_bf16x2(((mbits[c] >> (2 * j)) & 1u) ? v[2 * j] : 0.f, ((mbits[c] >> (2 * j + 1)) & 1u) ? v[2 * j + 1] : 0.f);
        } else if (p.relu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[c][j] = ptx::cvt_bf16x2_relu(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[c][j] = ptx::cvt_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        if (p.y_pool) {
          // 2x2 ceil-mode max of the bf16 results (post-ReLU, >= 0: an out-of-frame partner contributes 0)
          // (row first, then the row below: with `pool_arg` the window index of the FIRST maximum in scan order is kept --
          //  a later candidate wins only if strictly greater; non-negative bf16 compare as integers)
          uint32_t pl[8];
          uint32_t right_wins = 0u, row2_wins = 0u;          // bit 2 j + h: channel c0 + 2 j + h
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint32_t mine = valid ? o[c][j] : 0u;
            const uint32_t right = __shfl_down_sync(0xffffffffu, mine, 1);
            __nv_bfloat162 m = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&mine), *reinterpret_cast<const __nv_bfloat162*>(&right));
            const uint32_t am = *reinterpret_cast<const uint32_t*>(&m);
            const uint32_t below = __shfl_xor_sync(0xffffffffu, am, 16);
            m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(&below));
            pl[j] = *reinterpret_cast<const uint32_t*>(&m);
            if (p.pool_arg) {
              right_wins |= ((right & 0xffffu) > (mine & 0xffffu) ? 1u : 0u) << (2 * j) | ((right >> 16) > (mine >> 16) ? 1u : 0u) << (2 * j + 1);
              row2_wins |= ((below & 0xffffu) > (am & 0xffffu) ? 1u : 0u) << (2 * j) | ((below >> 16) > (am >> 16) ? 1u : 0u) << (2 * j + 1);
            }
          }
          if (pool_writer) ptx::stg256(p.y_pool + ppix * 64 + c0, pl);
          if (p.pool_arg) {
            const uint32_t right_wins_below = __shfl_xor_sync(0xffffffffu, right_wins, 16);
            arg_lo |= ((row2_wins & right_wins_below) | (~row2_wins & right_wins & 0xffffu)) << (16 * c);
            arg_hi |= row2_wins << (16 * c);
          }
        }
      }
      if (p.pool_arg && pool_writer)
        *reinterpret_cast<uint2*>(p.pool_arg + (ppix * 2 + (cbase >> 1)) * 2) = make_uint2(arg_lo, arg_hi);
      if (p.y) {
        // 2 x 2 transpose of 32-byte items inside the lane pair: o[c] (own pixel, chunk cbase + c) -> f[k] (pixel 2g + k, chunk cbase + j2)
        uint32_t f[2][8];
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const uint32_t recv = __shfl_xor_sync(0xffffffffu, j2 ? o[0][r] : o[1][r], 1);
          f[0][r] = j2 ? recv : o[0][r];
          f[1][r] = j2 ? o[1][r] : recv;
        }
#pragma unroll
        for (int k = 0; k < 2; ++k)
          if ((vmask >> (g2 + k)) & 1u) ptx::stg256(p.y + (pix - j2 + k) * 64 + (cbase + j2) * 16, f[k]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, SK_TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*SkEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SkEncodeTiledFn sk_get_encode() {
  static SkEncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<SkEncodeTiledFn>(f);
  else
    cudaGetLastError();
  return fn;
}

static void sk_fast_div(int d, uint32_t& mul, uint32_t& shift) {
  uint32_t l = 0;
  while ((1u << l) < (uint32_t)d) ++l;
  mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - (uint64_t)d)) / (uint64_t)d + 1);
  shift = l;
}

// does the row-stacked kernel serve this layer? (64 output channels, weights resident next to >= 3 halo boxes)
bool conv_stack_tc_supported(int Cin, int Cout) {
  if (Cout != 64 || Cin <= 0 || Cin % 64 != 0) return false;
  const int w_bytes = (Cin / 64) * 3 * SK_WT_BYTES;
  return (227 * 1024 - 1024 - SK_MISC_BYTES - w_bytes) / SK_A_BYTES >= 3;
}

// x (N,H,W,Cin) bf16, w_packed [64][9][Cin] bf16 (FOSVOS_W_TC_FWD / _DGRAD), y / mask (N,H,W,64) bf16, y_pool pooled.
int conv_stack_tc_launch(const void* x, const void* w_packed, const float* bias, const void* mask, void* y, void* y_pool, void* pool_arg,
                         int N, int H, int W, int Cin, int relu, cudaStream_t stream) {
  SkEncodeTiledFn enc = sk_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  SkParams p;
  p.bias = bias;
  p.mask = (const __nv_bfloat16*)mask;
  p.y = (__nv_bfloat16*)y;
  p.y_pool = (__nv_bfloat16*)y_pool;
  p.pool_arg = (uint32_t*)pool_arg;
  p.relu = relu;
  p.N = N; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, SK_OUT_W);
  p.tiles_y = ceil_div(H, SK_TH);
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  FOSVOS_REQUIRE(tiles < (1LL << 31), "conv3x3_tc (row stack): too many tiles");
  p.total_tiles = (int)tiles;
  p.k_chunks = Cin / 64;
  p.cin_pad = Cin;
  const int w_bytes = p.k_chunks * 3 * SK_WT_BYTES;
  p.stages = min(6, (227 * 1024 - 1024 - SK_MISC_BYTES - w_bytes) / SK_A_BYTES);
  sk_fast_div(p.tiles_x, p.dx_mul, p.dx_shift);
  sk_fast_div(p.tiles_y, p.dy_mul, p.dy_shift);

  CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, SK_TW, SK_TH + 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(row-stack activations %dx%dx%dx%d) failed: %d", N, H, W, Cin, (int)r); return FOSVOS_ERR_DRIVER; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * p.cin_pad, 64};
    cuuint64_t strides[1] = {(cuuint64_t)9 * p.cin_pad * 2};
    cuuint32_t box[2] = {64, 64};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(row-stack weights, cin_pad %d) failed: %d", p.cin_pad, (int)r); return FOSVOS_ERR_DRIVER; }
  }
  const int smem_bytes = p.stages * SK_A_BYTES + w_bytes + SK_MISC_BYTES + 1024;
  static unsigned long long attr_set = 0;          // one bit per device: function attributes are per device
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_stack_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { attr_set = 0; set_error("cudaFuncSetAttribute(row-stack smem): %s", cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
  }
  const int grid = min(p.total_tiles, num_sms());
  conv3x3_stack_tc_kernel<<<grid, SK_THREADS, smem_bytes, stream>>>(mx, mw, p);
  return check_launch("conv3x3_tc (row stack)");
}

}  // namespace fosvos
