// Weight gradient of the 3x3 convolution on the 5th-generation tensor cores.
//
//   dW[co][ci][tap] = sum_{pixels p} dZ[p][co] * X[p + tap][ci]            bf16 x bf16 -> fp32
//
// GEMM-K is the PIXEL dimension, so both operands are "MN-major": NHWC rows (one pixel = one
// 128-byte row of 64 channels) are exactly the SWIZZLE_128B MN-major canonical layout, and a
// 16-pixel K step is 16 consecutive rows.  One work item = (128-row M tile, N tile, kernel column
// s, pixel range); its three accumulators (kernel rows r = 0..2) live in TMEM side by side.
// Per 128-pixel patch (TH x TW, TW % 8 == 0) the CTA loads
//   - the un-shifted operand:  boxes {64ch, TW, TH, 1}
//   - the shifted operand X:   boxes {64ch, TW, TH+2, 1} at (x0+s-1, y0-1)  -- ONE halo box serves the
//     three vertical taps: tap r starts r*TW rows (a multiple of the 1024-byte swizzle atom) further.
// TMA zero-fills out-of-frame rows (= the convolution padding) and channel tails.
// Orientation is chosen per layer so that M = 128 is well filled:
//   X_IS_A = 0:  M = cout (dZ is A), N = cin tile      -> workspace [tap][cout][cin]
//   X_IS_A = 1:  M = cin  (X  is A), N = cout tile     -> workspace [tap][cin][cout]   (side_prep, N = 16)
// Split-K partial sums are merged with vectorised fp32 reductions (red.global.add.v4.f32) into a
// [tap][M][N] workspace; `wgrad_unpack_kernel` then adds it into the OIHW fp32 .grad tensor.  The
// workspace may stay live across micro-iterations (`_accumulate` / `_finish`), so the unpack runs once
// per optimizer step.
// First layer (Cin padded to 8, `c8`): a pixel of X is 16 bytes = one row of an UN-swizzled MN-major core matrix, and
// the frame is described to TMA as (8 W, H, N), so ONE 160 B x 18-row halo box per 16 x 8 patch serves all nine
// taps: tap (r, s) is the same tile read from byte offset r * 160 + s * 16.  With SBO = 16 (next tap along N) and
// LBO = 160 (next image row = next 8 pixels along K) one MMA covers the three taps of a kernel row (N = 32, the
// fourth block is ignored), so a patch is visited once (not once per kernel column) and dZ is read once.  And since a halo
// row is 10 such blocks long, block 10 r + s is tap (r, s): ONE N = 184 MMA per 16 pixels covers all nine taps (`c8_rows3`,
// 92 cycles instead of 3 x 40; the blocks in between multiply pixels further right and are never stored).
//
// Row stack (`stack`, N tile of 64 channels with X on N: conv2_1): the three vertical taps of the X halo box are three
// 64-column blocks of ONE N = 192 MMA -- an MN-major operand's 64-element blocks may sit anywhere LBO bytes apart, and
// LBO = one image row of the patch makes block r the box seen through kernel row r.  An M = 128 MMA costs
// max(N / 2, 32 + N / 4) cycles whatever its operands' major-ness (tools/exp/mma_side.cu, mma_major.cu: 48 / 64 / 96 / 128
// cycles for N = 64 / 128 / 192 / 256), so one N = 192 MMA (96 cycles, tensor-bound) replaces three N = 64 ones (144).
//
// X on M with a narrow dZ (`zstack`: side_prep, Cout = 16): the same stack with the vertical shift moved to dZ (below); its
// 16 channels sit in 64-column blocks (TMA zero-fills the rest), so a kernel column is one N = 192 MMA instead of three
// N = 16 ones (39 cycles each: a narrow MMA is bound by reading its 128-row A tile).
//
// Cout <= 64 with X on M (`one_pass`: conv1_2, and side_prep by default): the stack with the vertical shift moved to dZ -- sum_p dZ[p] X[p + (dy, dx)]
// = sum_q dZ[q - (dy, 0)] X[q + (0, dx)] -- so B = dZ halo box (three row shifts, N = 192) and A = X under two horizontal
// shifts (two plain boxes LBO apart, M = 128).  Two MMAs per 16 pixels -- kernel columns (0, 1) and (2, ignored rows) --
// cover all nine taps from ONE visit of the patch (3 X boxes + 1 dZ halo box = 68 KB), instead of six N = 64 MMAs in
// each of two passes over both operands.  Wider X is tiled by 64 channels (one work item each).
//
// Bias gradient db[co] = sum_p dZ[p][co]: the dZ boxes are in shared memory anyway, so the four epilogue
// warps, otherwise idle until the accumulators are complete, sum them on the side; the CTAs that see the
// same dZ tile (three kernel columns x the tiles of the X-channel dimension) take every share-th patch each.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace fosvos {

constexpr int WG_THREADS = 192;
constexpr int WG_PLAIN_BLOCK = 128 * 128;          // bytes of one un-shifted 64-channel box (128 px)

struct WgParams {
  float* ws;                 // [9][Mtot][Ntot] fp32
  float* db;                 // bias gradient (Cout fp32, accumulated) or null
  int cout;                  // real output channels (db length)
  int N, H, W;
  int Mtot, Ntot;            // padded channel counts of the M and N dimensions
  int m_tiles, n_tiles, splits;
  int n_cols;                // N extent per tap (16, 64 or 128)
  int nb_n;                  // 64-channel boxes per N tile
  int tiles_x, tiles_y, patches;
  int tw_shift;              // TW = 1 << tw_shift (>= 8)
  int x_is_a;
  int stages, stage_bytes, a_bytes;
  int x_block;               // bytes of one shifted 64-channel box: (TH+2)*TW*128
  int c8;                    // first-layer mode (see the header comment)
  int stack;                 // N = 64 tile, X on N: the three kernel rows are one N = 192 MMA (LBO = one image row of the patch)
  int one_pass;              // Cin == Cout == 64: A = X under two column shifts, B = dZ under three row shifts, all taps per visit
  int zstack;                // X on M, dZ (<= 64 channels) as ONE halo box whose three row shifts are one N = 192 MMA (one_pass, side_prep)
  int c8_lbo, c8_sbo;        // descriptor strides of its un-swizzled X operand (160 / 16)
  int c8_rows3;              // first layer: the three kernel rows in one N = 184 MMA
  int debug;                 // FOSVOS_WG_DEBUG (timing experiments only): 1 = skip the reductions, 2 = no start rotation, 3 = no MMAs, 4 = no loads, 5 = neither, 6 = 5 + 1
};

// MN-major SWIZZLE_128B operand: rows (K) of 128 B, 8-row groups SBO = 1024 B apart, 64-element
// MN blocks LBO bytes apart.
__device__ __forceinline__ uint64_t umma_desc_sw128_mnmajor(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

constexpr int WG_C8_ROW = 160, WG_C8_BOX = 18 * WG_C8_ROW, WG_C8_BYTES = 3072;

// Un-swizzled MN-major operand: 16 contiguous bytes along MN, 8 K rows 16 bytes apart form a core matrix;
// `sbo` = byte distance between core matrices along MN, `lbo` = along K.
__device__ __forceinline__ uint64_t umma_desc_noswizzle_mnmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z, const WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* done_bar = empty_bar + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  __shared__ float bias_red[128];

  // work item
  int item = blockIdx.x;
  const int split = item % p.splits; item /= p.splits;
  const int n_s = (p.c8 || p.one_pass) ? 1 : 3;     // passes over the kernel columns
  const int s = item % n_s;                         // kernel column of this pass
  item /= n_s;
  const int nt = item % p.n_tiles;
  const int mt = item / p.n_tiles;
  // Bias gradient: the CTAs (s, other-dimension tile) that share this CTA's dZ tile split its patches between them
  // (patch pt belongs to CTA pt % share == share_id), so the extra shared-memory reads are spread evenly instead of
  // slowing one CTA in three -- the MMA stream already uses the full shared-memory bandwidth.
  const bool do_bias = p.db != nullptr;
  const int share = n_s * (p.x_is_a ? p.m_tiles : p.n_tiles);
  const int share_id = s * (p.x_is_a ? p.m_tiles : p.n_tiles) + (p.x_is_a ? mt : nt);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform roles
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_z);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], do_bias ? 5 : 1);      // MMA commit (+ the four bias-summing warps)
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 128) bias_red[threadIdx.x] = 0.f;
  if (p.c8) {
    // tails of the X stages: never written by TMA, read (and ignored) by the fourth column block of the last kernel row
    const int per = (WG_C8_BYTES - WG_C8_BOX) / 16;
    for (int i = threadIdx.x; i < p.stages * per; i += WG_THREADS)
      *reinterpret_cast<uint4*>(smem + (i / per) * p.stage_bytes + p.a_bytes + WG_C8_BOX + (i % per) * 16) = make_uint4(0u, 0u, 0u, 0u);
    ptx::fence_proxy_async_smem();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int p_begin = (int)((long long)p.patches * split / p.splits);
  const int p_end = (int)((long long)p.patches * (split + 1) / p.splits);
  const int TW = 1 << p.tw_shift, TH = 128 >> p.tw_shift;
  const int m0 = mt * 128, n0 = nt * p.n_cols;
  // channel origins of the two operands
  const int xc0 = p.x_is_a ? m0 : n0;          // X channels (cin)
  const int zc0 = p.x_is_a ? n0 : m0;          // dZ channels (cout)
  const int nb_x = p.x_is_a ? 2 : p.nb_n;
  const int nb_z = p.x_is_a ? p.nb_n : 2;
  const int x_off = p.x_is_a ? 0 : p.a_bytes;  // byte offset of the X boxes inside a stage
  const int z_off = p.zstack ? p.a_bytes + TW * 128 : p.x_is_a ? p.a_bytes : 0;   // (zstack: the patch inside the dZ halo box)

  if (warp == 0) {
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = p_begin; pt < p_end; ++pt) {
        int t = pt;
        const int tx = t % p.tiles_x; t /= p.tiles_x;
        const int ty = t % p.tiles_y;
        const int n = t / p.tiles_y;
        const int x0 = tx * TW, y0 = ty * TH;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* st = smem + stage * p.stage_bytes;
        if (p.c8) {
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full_bar[stage], WG_C8_BOX + nb_z * WG_PLAIN_BLOCK);
            ptx::tma_load_3d_a(ptx::smem_u32(st + x_off), &map_x, ptx::smem_u32(&full_bar[stage]), (x0 - 1) * 8, y0 - 1, n);
            for (int j = 0; j < nb_z; ++j)
              ptx::tma_load_4d(st + z_off + j * WG_PLAIN_BLOCK, &map_z, &full_bar[stage], zc0 + 64 * j, x0, y0, n);
          }
        } else if (p.debug >= 4) {              // timing experiment: no loads
          if (ptx::elect_one()) ptx::mbar_arrive(&full_bar[stage]);
        } else if (p.zstack) {
          if (ptx::elect_one()) {
            ptx::mbar_expect_tx(&full_bar[stage], p.stage_bytes);
            // plain X boxes: one_pass = the 64 channels through kernel columns 0..2; else the tile's two channel blocks through column s
            for (int j = 0; j < (p.one_pass ? 3 : 2); ++j)
              ptx::tma_load_4d(st + j * WG_PLAIN_BLOCK, &map_x, &full_bar[stage], p.one_pass ? mt * 64 : xc0 + 64 * j, x0 + (p.one_pass ? j : s) - 1, y0, n);
            ptx::tma_load_4d(st + p.a_bytes, &map_z, &full_bar[stage], 0, x0, y0 - 1, n);      // the dZ halo box
          }
        } else if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[stage], p.stage_bytes);
          for (int j = 0; j < nb_x; ++j)
            ptx::tma_load_4d(st + x_off + j * p.x_block, &map_x, &full_bar[stage], xc0 + 64 * j, x0 + s - 1, y0 - 1, n);
          for (int j = 0; j < nb_z; ++j)
            ptx::tma_load_4d(st + z_off + j * WG_PLAIN_BLOCK, &map_z, &full_bar[stage], zc0 + 64 * j, x0, y0, n);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      // kind::f16, D fp32, A/B bf16, both MN-major (bits 15/16), M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) |
                             ((uint32_t)(p.n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t lbo_a = (p.x_is_a && !p.zstack) ? p.x_block : WG_PLAIN_BLOCK;
      const uint32_t lbo_b = p.x_is_a ? WG_PLAIN_BLOCK : p.x_block;
      const uint32_t tap_bytes = (uint32_t)TW * 128;       // one image row of the patch
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = p_begin; pt < p_end; ++pt) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + stage * p.stage_bytes);
        const uint32_t a_base = st, b_base = st + p.a_bytes;
        if (p.c8) {
          if (ptx::elect_one()) {
            if (p.c8_rows3) {
              // all nine taps in ONE MMA per 16 pixels: the 8-channel blocks of B are 16 B apart along N, and a halo row is
              // 10 blocks long, so block 10 r + s is tap (r, s): N = 8 * 23 = 184 (92 cycles) instead of three N = 32
              // MMAs (3 x 40); the blocks in between multiply pixels further right and are never stored
              const uint32_t idesc184 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(184 >> 3) << 17);
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t da = umma_desc_sw128_mnmajor(a_base + kk * 2048, WG_PLAIN_BLOCK);
                const uint64_t db = umma_desc_noswizzle_mnmajor(b_base + 2 * kk * WG_C8_ROW, p.c8_lbo, p.c8_sbo);
                ptx::umma_bf16(tmem_base, da, db, idesc184, (pt != p_begin) || (kk != 0));
              }
            } else {
#pragma unroll 1
            for (int r = 0; r < 3; ++r) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const uint64_t da = umma_desc_sw128_mnmajor(a_base + kk * 2048, WG_PLAIN_BLOCK);
                // 16 pixels = image rows 2 kk, 2 kk + 1 of the patch, seen through kernel row r
                const uint64_t db = umma_desc_noswizzle_mnmajor(b_base + (r + 2 * kk) * WG_C8_ROW, p.c8_lbo, p.c8_sbo);
                ptx::umma_bf16(tmem_base + r * p.n_cols, da, db, idesc, (pt != p_begin) || (kk != 0));
              }
            }
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (pt == p_end - 1) ptx::umma_commit(done_bar);
          }
        } else if (p.zstack || p.stack) {
          if (ptx::elect_one()) {
            const uint32_t idesc192 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(192 >> 3) << 17);
            if (p.debug != 3 && p.debug < 5) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                // B: three 64-column blocks one image row of the patch apart = the box through kernel rows 0..2
                const uint64_t db = umma_desc_sw128_mnmajor(b_base + kk * 2048, tap_bytes);
                const uint32_t acc = (pt != p_begin) || (kk != 0);
                ptx::umma_bf16(tmem_base, umma_desc_sw128_mnmajor(a_base + kk * 2048, lbo_a), db, idesc192, acc);
                // one_pass: rows 0..63 = kernel column 2, rows 64..127 read the head of the dZ box (never stored)
                if (p.one_pass) ptx::umma_bf16(tmem_base + 192, umma_desc_sw128_mnmajor(a_base + 2 * WG_PLAIN_BLOCK + kk * 2048, lbo_a), db, idesc192, acc);
              }
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (pt == p_end - 1) ptx::umma_commit(done_bar);
          }
        } else if (ptx::elect_one()) {
#pragma unroll 1
        for (int r = 0; r < ((p.debug == 3 || p.debug >= 5) ? 0 : 3); ++r) {        // debug 3: timing experiment without the MMAs
          const uint32_t a_r = a_base + (p.x_is_a ? r * tap_bytes : 0);
          const uint32_t b_r = b_base + (p.x_is_a ? 0 : r * tap_bytes);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t da = umma_desc_sw128_mnmajor(a_r + kk * 2048, lbo_a);
            const uint64_t db = umma_desc_sw128_mnmajor(b_r + kk * 2048, lbo_b);
            ptx::umma_bf16(tmem_base + r * p.n_cols, da, db, idesc, (pt != p_begin) || (kk != 0));
          }
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (pt == p_end - 1) ptx::umma_commit(done_bar);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue: TMEM -> vectorised fp32 reductions =====================
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    if (do_bias) {
      // thread t of the 128: 16-byte chunk c (8 channels) of pixel rows rg, rg + 16, ... of every dZ box.
      // SWIZZLE_128B puts chunk c of row r at r * 128 + ((c ^ (r & 7)) << 4); r & 7 == rg & 7 for all its rows.
      const int t = threadIdx.x - 64;
      const int c = t & 7, rg = t >> 3;
      const uint32_t off = (uint32_t)rg * 128 + (uint32_t)((c ^ (rg & 7)) << 4);
      float acc[2][8];
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[b][e] = 0.f;
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = p_begin; pt < p_end; ++pt) {
        ptx::mbar_wait(&full_bar[stage], phase);
        const uint8_t* zb = smem + stage * p.stage_bytes + z_off + off;
        const bool mine = (pt % share) == share_id;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          if (mine && b < nb_z) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = *reinterpret_cast<const uint4*>(zb + b * WG_PLAIN_BLOCK + j * 2048);
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[b][2 * q] += __uint_as_float(w[q] << 16);
                acc[b][2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&bias_red[b * 64 + c * 8 + e], acc[b][e]);
    }
    if (p_end > p_begin) {
      ptx::mbar_wait_idle(done_bar, 0);
      ptx::tc_fence_after();
      const bool row_ok = (m0 + row) < p.Mtot && p.debug != 1 && p.debug != 6;
      // the splits of one tile finish together and add into the same addresses: start each at a different
      // (kernel row, column block) so that concurrent reductions mostly hit different L2 lines
      const int rot = p.debug == 2 ? 0 : split;
      const int n_cb = p.n_cols >> 4;
      if (p.c8) {
        // accumulator r, column block sc = tap (r, sc), 8 channels each: ws[tap][cout][8]
#pragma unroll 1
        for (int r = 0; r < 3; ++r) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + r * (p.c8_rows3 ? 80 : p.n_cols);
#pragma unroll 1
          for (int c0 = 0; c0 < 32; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(taddr + c0, v);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int sc = (c0 >> 3) + h;
                if (sc < 3) {
                  float* dst = p.ws + ((long long)(r * 3 + sc) * p.Mtot + (m0 + row)) * 8;
                  red_add_v4(dst, __uint_as_float(v[8 * h]), __uint_as_float(v[8 * h + 1]), __uint_as_float(v[8 * h + 2]), __uint_as_float(v[8 * h + 3]));
                  red_add_v4(dst + 4, __uint_as_float(v[8 * h + 4]), __uint_as_float(v[8 * h + 5]), __uint_as_float(v[8 * h + 6]), __uint_as_float(v[8 * h + 7]));
                }
              }
            }
          }
        }
      } else if (p.one_pass) {
        // accumulator a (0: kernel columns 0 / 1 on rows 0..63 / 64..127; 1: column 2 on rows 0..63), column block i = the dZ
        // box shifted down by i rows = kernel row 2 - i: ws[tap][cin][cout]
        for (int ai = 0; ai < 6; ++ai) {
          const int a = ((ai + rot) % 6) / 3, i = (ai + rot) % 3;
          const int sj = a ? 2 : (row >> 6);
          const int ci = mt * 64 + (row & 63);              // (one_pass tiles the X channels by 64: m_tiles = CinP / 64)
          const bool ok = p.debug != 1 && p.debug != 6 && (a == 0 || row < 64) && ci < p.Mtot;
          float* dst = p.ws + ((long long)((2 - i) * 3 + sj) * p.Mtot + ci) * p.Ntot;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + a * 192 + i * 64;
#pragma unroll 1
          for (int c0 = 0; c0 < p.Ntot; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(taddr + c0, v);
            ptx::tmem_ld_wait();
            if (ok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (c0 + 4 * q < p.Ntot)
                  red_add_v4(dst + c0 + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                             __uint_as_float(v[4 * q + 3]));
              }
            }
          }
        }
      } else if (p.zstack) {
        // column block i = the dZ box shifted down by i rows = kernel row 2 - i; the Ntot (<= 64) real channels lead each block
        for (int ii = 0; ii < 3; ++ii) {
          const int i = (ii + rot) % 3;
          float* dst = p.ws + ((long long)((2 - i) * 3 + s) * p.Mtot + m0 + row) * p.Ntot;
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + i * 64;
#pragma unroll 1
          for (int c0 = 0; c0 < p.Ntot; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(taddr + c0, v);
            ptx::tmem_ld_wait();
            if (row_ok) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                if (c0 + 4 * q < p.Ntot)
                  red_add_v4(dst + c0 + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                             __uint_as_float(v[4 * q + 3]));
              }
            }
          }
        }
      } else
      for (int rr = 0; rr < 3; ++rr) {
        const int r = (rr + rot) % 3;
        const int tap = r * 3 + s;
        const int mrow = m0 + row;
        float* dst = p.ws + ((long long)tap * p.Mtot + mrow) * p.Ntot + n0;
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + r * p.n_cols;
#pragma unroll 1
        for (int cb = 0; cb < n_cb; ++cb) {
          const int c0 = ((cb + rot / 3) % n_cb) << 4;
          uint32_t v[16];
          ptx::tmem_ld16(taddr + c0, v);
          ptx::tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (n0 + c0 + 4 * q < p.Ntot)
                red_add_v4(dst + c0 + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                           __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            }
          }
        }
      }
    }
    if (do_bias) {
      // last, off the critical path: this CTA's share of the bias gradient (its shared-memory partial sums are complete)
      ptx::named_bar_sync(1, 128);
      const int t = threadIdx.x - 64;
      if (zc0 + t < p.cout && t < 64 * nb_z) atomicAdd(p.db + zc0 + t, bias_red[t]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) for the wide layers: Cout % 256 == 0, Cin % 128 == 0, dZ on M.
//
// A single-CTA M = 128, N = 128 MMA is exactly operand-read balanced (A 4 KB + B 4 KB at 128 B/clk = 64 cycles for 64
// cycles of tensor work), so every stall shows, and each CTA pulls 72 KB per 128-pixel patch through L2.  A pair of CTAs
// on the two SMs of a TPC runs ONE M = 256 x N = 128 MMA per step: each SM reads its own 128 dZ channels (A) and only
// HALF of the X tile (64 of the 128 cin channels; the tensor cores exchange the halves), 48 operand cycles under 64 tensor
// cycles, and 52 KB per patch and CTA through L2.  Work item = (256-cout tile, 128-cin tile, kernel column s, pixel
// range); the three accumulators (kernel rows) are 128 columns each in BOTH CTAs' TMEM (rows 0..127 / 128..255 of the
// cout tile).  Both CTAs run a TMA producer for their own halves; all bytes of a stage are counted on CTA 0's `full`
// barrier, where the one MMA-issuing thread of the pair waits; its commits arrive on both CTAs' `empty` / `done`
// barriers (multicast).  CTA 1's bias-summing warps learn that a stage has landed from a remote arrive on `landed`.
struct WgPairParams {
  float* ws;                 // [9][CoutP][CinP] fp32
  float* db;
  int cout;
  int CoutP, CinP;
  int n_tiles, splits;       // 128-cin tiles; pixel-range splits
  int tiles_x, tiles_y, patches;
  int tw_shift;
  int stages, stage_bytes, x_block;
  int debug;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WG_THREADS, 1)
conv3x3_wgrad_tc_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_z, const WgPairParams p) {
  extern __shared__ uint8_t smem_raw[];
  // both CTAs must lay their shared memory out identically: the MMA descriptors and the multicast commits address by offset
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.stages * p.stage_bytes);
  uint64_t* empty_bar = full_bar + p.stages;
  uint64_t* landed_bar = empty_bar + p.stages;
  uint64_t* done_bar = landed_bar + p.stages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(done_bar + 1);
  __shared__ float bias_red[128];

  const uint32_t rank = ptx::cluster_ctarank();
  int item = blockIdx.x >> 1;
  const int split = item % p.splits; item /= p.splits;
  const int s = item % 3; item /= 3;
  const int nt = item % p.n_tiles;
  const int mt = item / p.n_tiles;
  const bool do_bias = p.db != nullptr;
  const int share = 3 * p.n_tiles;                 // CTAs that see this CTA's dZ half-tile
  const int share_id = s * p.n_tiles + nt;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_z);
    for (int i = 0; i < p.stages; ++i) {
      ptx::mbar_init(&full_bar[i], 1);                      // CTA 0's producer (expect_tx for both CTAs' bytes)
      ptx::mbar_init(&empty_bar[i], do_bias ? 5 : 1);       // the pair's MMA commit (+ this CTA's four bias-summing warps)
      ptx::mbar_init(&landed_bar[i], 1);                    // CTA 1 only: remote arrive from CTA 0's MMA warp
    }
    ptx::mbar_init(done_bar, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 128) bias_red[threadIdx.x] = 0.f;
  if (warp == 1) ptx::tmem_alloc_pair(tmem_ptr, 512);
  ptx::tc_fence_before();
  ptx::cluster_sync();                                      // barriers of both CTAs initialised before any remote traffic
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int p_begin = (int)((long long)p.patches * split / p.splits);
  const int p_end = (int)((long long)p.patches * (split + 1) / p.splits);
  const int TW = 1 << p.tw_shift, TH = 128 >> p.tw_shift;
  const int zc0 = mt * 256 + (int)rank * 128;      // this CTA's dZ channels (its 128 rows of the M = 256 tile)
  const int xc0 = nt * 128 + (int)rank * 64;       // this CTA's half of the X tile (64 of the N = 128 columns)
  const int a_bytes = 2 * WG_PLAIN_BLOCK;

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int pt = p_begin; pt < p_end; ++pt) {
      int t = pt;
      const int tx = t % p.tiles_x; t /= p.tiles_x;
      const int ty = t % p.tiles_y;
      const int n = t / p.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH;
      ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
      uint8_t* st = smem + stage * p.stage_bytes;
      if (ptx::elect_one()) {
        const uint32_t lead = ptx::mapa(ptx::smem_u32(&full_bar[stage]), 0);
        if (p.debug >= 4) {                     // timing experiment: no loads, the MMAs run on whatever the stage holds
          if (rank == 0) ptx::mbar_arrive(&full_bar[stage]);
        } else {
        if (rank == 0) ptx::mbar_expect_tx(&full_bar[stage], 2 * p.stage_bytes);
        ptx::tma_load_4d_pair(st, &map_z, lead, zc0, x0, y0, n);
        ptx::tma_load_4d_pair(st + WG_PLAIN_BLOCK, &map_z, lead, zc0 + 64, x0, y0, n);
        ptx::tma_load_4d_pair(st + a_bytes, &map_x, lead, xc0, x0 + s - 1, y0 - 1, n);
        }
      }
      __syncwarp();
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // kind::f16, D fp32, A/B bf16, both MN-major (bits 15/16), M = 256 over the pair, N = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint32_t tap_bytes = (uint32_t)TW * 128;
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = p_begin; pt < p_end; ++pt) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(smem + stage * p.stage_bytes);
        if (do_bias && lane == 31) ptx::mbar_arrive_cluster_relaxed(ptx::mapa(ptx::smem_u32(&landed_bar[stage]), 1));
        if (ptx::elect_one()) {
#pragma unroll 1
          for (int r = 0; r < ((p.debug == 3 || p.debug >= 5) ? 0 : 3); ++r) {      // debug 3: timing experiment without the MMAs
            const uint32_t b_r = st + a_bytes + r * tap_bytes;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const uint64_t da = umma_desc_sw128_mnmajor(st + kk * 2048, WG_PLAIN_BLOCK);
              const uint64_t db = umma_desc_sw128_mnmajor(b_r + kk * 2048, p.x_block);
              ptx::umma_bf16_pair(tmem_base + r * 128, da, db, idesc, (pt != p_begin) || (kk != 0));
            }
          }
          ptx::umma_commit_pair(&empty_bar[stage], 3);
          if (pt == p_end - 1) ptx::umma_commit_pair(done_bar, 3);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    if (do_bias) {
      const int t = threadIdx.x - 64;
      const int c = t & 7, rg = t >> 3;
      const uint32_t off = (uint32_t)rg * 128 + (uint32_t)((c ^ (rg & 7)) << 4);
      float acc[2][8];
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[b][e] = 0.f;
      uint64_t* ready = rank == 0 ? full_bar : landed_bar;
      int stage = 0;
      uint32_t phase = 0;
      for (int pt = p_begin; pt < p_end; ++pt) {
        ptx::mbar_wait(&ready[stage], phase);
        const uint8_t* zb = smem + stage * p.stage_bytes + off;
        if ((pt % share) == share_id) {
#pragma unroll
          for (int b = 0; b < 2; ++b) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint4 v = *reinterpret_cast<const uint4*>(zb + b * WG_PLAIN_BLOCK + j * 2048);
              const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                acc[b][2 * q] += __uint_as_float(w[q] << 16);
                acc[b][2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int e = 0; e < 8; ++e) atomicAdd(&bias_red[b * 64 + c * 8 + e], acc[b][e]);
    }
    ptx::mbar_wait_idle(done_bar, 0);
    ptx::tc_fence_after();
    const int rot = p.debug == 2 ? 0 : split;
    for (int rr = 0; rr < 3; ++rr) {
      const int r = (rr + rot) % 3;
      float* dst = p.ws + ((long long)(r * 3 + s) * p.CoutP + zc0 + row) * p.CinP + nt * 128;
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + r * 128;
#pragma unroll 1
      for (int cb = 0; cb < 8; ++cb) {
        const int c0 = ((cb + rot / 3) & 7) << 4;
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tmem_ld_wait();
        if (p.debug != 1 && p.debug != 6) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            red_add_v4(dst + c0 + 4 * q, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                       __uint_as_float(v[4 * q + 3]));
        }
      }
    }
    if (do_bias) {
      ptx::named_bar_sync(1, 128);
      const int t = threadIdx.x - 64;
      if (zc0 + t < p.cout) atomicAdd(p.db + zc0 + t, bias_red[t]);
    }
  }
  // neither CTA may leave (or free TMEM) while the other's MMAs / commits / TMA signals can still touch it
  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

// dw_oihw[co][ci][tap] += ws[tap][a][b];  ws is [tap][cout][cin] (x_is_a = 0) or [tap][cin][cout]
// (padded rows / columns of ws only ever receive zeros, so they need no clearing)
__global__ void wgrad_unpack_kernel(float* __restrict__ ws, float* __restrict__ dw, int Cout, int Cin, int CoutP,
                                    int CinP, int x_is_a, int zero_ws) {
  const int total = Cout * Cin;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    // fastest index follows the workspace's contiguous dimension -> coalesced reads
    int co, ci;
    if (x_is_a) { co = i % Cout; ci = i / Cout; } else { ci = i % Cin; co = i / Cin; }
    const long long plane = (long long)CoutP * CinP;
    const long long src = x_is_a ? ((long long)ci * CoutP + co) : ((long long)co * CinP + ci);
    float* d = dw + ((long long)co * Cin + ci) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      d[t] += ws[t * plane + src];
      if (zero_ws) ws[t * plane + src] = 0.f;
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn wg_get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    else
      cudaGetLastError();
  });
  return fn;
}

static int wg_encode(CUtensorMap* m, const void* x, int N, int H, int W, int C, int TW, int rows) {
  EncodeTiledFn enc = wg_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)TW, (cuuint32_t)rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad %dx%dx%dx%d box %dx%d) failed: %d", N, H, W, C, rows, TW, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

// first layer: the (N,H,W,8) frame as (8 W, H, N); box = one 16 x 8 patch with its halo, 10 pixels x 18 rows, no swizzle
static int wg_encode_c8(CUtensorMap* m, const void* x, int N, int H, int W) {
  EncodeTiledFn enc = wg_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {80, 18, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(wgrad first layer %dx%dx%d) failed: %d", N, H, W, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

size_t fosvos_conv3x3_wgrad_tc_workspace_bytes(int CinP, int CoutP) { return (size_t)9 * CinP * CoutP * sizeof(float); }

// orientation: put the wider channel dimension on M (=128 rows); tiny couts (side_prep) go to N
// (Cin = Cout = 64 goes to the X-on-M orientation for its two-columns-per-pass mode)
static inline int wg_x_is_a(int CinP, int CoutP) { return (CoutP < 64 || (CinP >= 128 && CoutP < 128) || (CinP == 64 && CoutP == 64)) ? 1 : 0; }

int fosvos_conv3x3_wgrad_tc_orientation(int CinP, int CoutP) { return wg_x_is_a(CinP, CoutP); }

int fosvos_conv3x3_wgrad_tc_accumulate(const void* x, const void* dz, float* db, void* workspace, int N, int H, int W,
                                       int CinP, int CoutP, int Cout, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && dz && workspace && N > 0 && H > 0 && W > 0, "conv3x3_wgrad_tc: null pointer or empty shape");
  FOSVOS_REQUIRE(CinP % 8 == 0 && CoutP % 8 == 0 && CinP > 0 && CoutP > 0 && Cout > 0 && Cout <= CoutP,
                 "conv3x3_wgrad_tc: bad channel counts (CinP=%d CoutP=%d Cout=%d)", CinP, CoutP, Cout);
  FOSVOS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)dz & 15) == 0 && ((uintptr_t)workspace & 15) == 0,
                 "conv3x3_wgrad_tc: pointers must be 16-byte aligned");
  cudaStream_t st = as_stream(stream);
  WgParams p;
  p.ws = (float*)workspace;
  p.db = db;
  p.cout = Cout;
  p.N = N; p.H = H; p.W = W;
  p.x_is_a = wg_x_is_a(CinP, CoutP);
  p.Mtot = p.x_is_a ? CinP : CoutP;
  p.Ntot = p.x_is_a ? CoutP : CinP;
  p.c8 = (CinP == 8 && !p.x_is_a && !getenv("FOSVOS_WG_NO_C8")) ? 1 : 0;
  static const bool no_stack = getenv("FOSVOS_WG_NO_STACK") != nullptr;      // A/B switch (timing experiments)
  // (narrow Cout, side_prep: 170 -> 162 us per window against three `zstack` passes; FOSVOS_WG_SIDE_THREE_PASSES selects those)
  static const bool side_one_pass = getenv("FOSVOS_WG_SIDE_THREE_PASSES") == nullptr;
  p.one_pass = (p.x_is_a && !no_stack && ((CinP == 64 && CoutP == 64) || (side_one_pass && CoutP <= 64 && CoutP % 4 == 0))) ? 1 : 0;
  p.c8_lbo = WG_C8_ROW; p.c8_sbo = 16;
  p.c8_rows3 = (p.c8 && !getenv("FOSVOS_WG_C8_THREE_MMAS")) ? 1 : 0;
  p.n_cols = p.c8 ? 32 : p.Ntot >= 128 ? 128 : (p.Ntot > 16 ? 64 : 16);
  p.nb_n = (p.n_cols + 63) / 64;
  p.stack = (!p.x_is_a && !p.c8 && p.Ntot == 64 && !no_stack) ? 1 : 0;
  p.zstack = (p.one_pass || (p.x_is_a && p.Ntot <= 64 && p.Ntot % 4 == 0 && !no_stack)) ? 1 : 0;
  p.m_tiles = ceil_div(p.Mtot, p.one_pass ? 64 : 128);
  p.n_tiles = ceil_div(p.Ntot, p.n_cols);
  // 128-pixel patch with TW % 8 == 0 that wastes the fewest out-of-frame pixels
  int best = 3; long long best_area = -1;
  for (int sh = 3; sh <= 6; ++sh) {
    const int TW = 1 << sh, TH = 128 >> sh;
    const long long area = (long long)ceil_div(H, TH) * TH * ceil_div(W, TW) * TW;
    if (best_area < 0 || area < best_area) { best_area = area; best = sh; }
  }
  if (p.c8) best = 3;                                 // 16 x 8 patches: the halo box layout is built for TW = 8
  p.tw_shift = best;
  const int TW = 1 << best, TH = 128 >> best;
  p.tiles_x = ceil_div(W, TW);
  p.tiles_y = ceil_div(H, TH);
  p.patches = N * p.tiles_x * p.tiles_y;
  p.x_block = (TH + 2) * TW * 128;
  const int nb_x = p.x_is_a ? 2 : p.nb_n, nb_z = p.x_is_a ? p.nb_n : 2;
  // zstack: X as plain boxes, dZ as ONE halo box (kernel rows)
  const int x_bytes = p.c8 ? WG_C8_BYTES : p.zstack ? (p.one_pass ? 3 : 2) * WG_PLAIN_BLOCK : nb_x * p.x_block;
  const int z_bytes = p.zstack ? p.x_block : nb_z * WG_PLAIN_BLOCK;
  p.a_bytes = p.x_is_a ? x_bytes : z_bytes;
  p.stage_bytes = x_bytes + z_bytes;
  p.stages = min(6, (220 * 1024 - 2048) / p.stage_bytes);
  FOSVOS_REQUIRE(p.stages >= 2, "conv3x3_wgrad_tc: stage of %d bytes does not fit twice in shared memory", p.stage_bytes);
  // Split-K over pixel ranges.  Every CTA ends with 3 * 128 * n_cols fp32 reductions into the workspace, so the
  // split count trades tensor-core occupancy against reduction traffic: fill the machine once; go to a second
  // wave only while each CTA still has enough patches to amortise its epilogue.
  const int items = p.m_tiles * p.n_tiles * ((p.c8 || p.one_pass) ? 1 : 3);
  int splits = max(1, num_sms() / items);
  static const bool two_waves = getenv("FOSVOS_WG_TWO_WAVES") != nullptr;   // (measured slower: 881 -> 822 us over the window's layers)
  if (two_waves && (long long)p.patches >= 16LL * 2 * num_sms() / items) splits = max(1, (2 * num_sms()) / items);
  if (const char* e = getenv("FOSVOS_WG_SPLITS")) { const int v = atoi(e); if (v > 0) splits = v; }
  splits = min(splits, p.patches);
  p.splits = splits;
  { const char* e = getenv("FOSVOS_WG_DEBUG"); p.debug = e ? atoi(e) : 0; }

  // wide layers: CTA-pair kernel (M = 256 over two SMs)
  static const bool pair_off = getenv("FOSVOS_WG_NO_CTA_PAIR") != nullptr;
  if (!pair_off && !p.x_is_a && !p.c8 && CoutP % 256 == 0 && CinP % 128 == 0) {
    WgPairParams q;
    q.ws = p.ws; q.db = db; q.cout = Cout; q.CoutP = CoutP; q.CinP = CinP;
    q.n_tiles = CinP / 128;
    q.tiles_x = p.tiles_x; q.tiles_y = p.tiles_y; q.patches = p.patches; q.tw_shift = p.tw_shift;
    q.x_block = p.x_block;
    q.stage_bytes = 2 * WG_PLAIN_BLOCK + p.x_block;
    q.stages = min(6, (220 * 1024 - 2048) / q.stage_bytes);
    q.debug = p.debug;
    const int pair_items = (CoutP / 256) * q.n_tiles * 3;
    int psplits = max(1, (num_sms() / 2) / pair_items);        // one wave of CTA pairs: one epilogue per SM
    if (const char* e = getenv("FOSVOS_WG_SPLITS")) { const int v = atoi(e); if (v > 0) psplits = v; }
    q.splits = min(psplits, p.patches);
    CUtensorMap mx, mz;
    int rc = wg_encode(&mx, x, N, H, W, CinP, TW, TH + 2);
    if (rc) return rc;
    rc = wg_encode(&mz, dz, N, H, W, CoutP, TW, TH);
    if (rc) return rc;
    static unsigned long long pair_smem_set = 0;
    if (first_use_on_device(pair_smem_set)) {
      cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tc_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
      if (e != cudaSuccess) { pair_smem_set = 0; set_error("cudaFuncSetAttribute(wgrad pair smem): %s", cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
    }
    conv3x3_wgrad_tc_pair_kernel<<<2 * pair_items * q.splits, WG_THREADS, q.stages * q.stage_bytes + 1024 + 1024, st>>>(mx, mz, q);
    return check_launch("conv3x3_wgrad_tc(pair)");
  }

  CUtensorMap mx, mz;
  int rc = p.c8 ? wg_encode_c8(&mx, x, N, H, W) : wg_encode(&mx, x, N, H, W, CinP, TW, p.zstack ? TH : TH + 2);
  if (rc) return rc;
  rc = wg_encode(&mz, dz, N, H, W, CoutP, TW, p.zstack ? TH + 2 : TH);
  if (rc) return rc;

  const int smem_bytes = p.stages * p.stage_bytes + 1024 + 1024;
  static unsigned long long smem_set = 0;          // one bit per device: function attributes are per device
  if (first_use_on_device(smem_set)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e != cudaSuccess) { smem_set = 0; set_error("cudaFuncSetAttribute(wgrad smem): %s", cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
  }
  conv3x3_wgrad_tc_kernel<<<items * splits, WG_THREADS, smem_bytes, st>>>(mx, mz, p);
  return check_launch("conv3x3_wgrad_tc");
}

int fosvos_conv3x3_wgrad_tc_finish(void* workspace, float* dw, int CinP, int CoutP, int Cin, int Cout, int zero_workspace,
                                   fosvos_stream_t stream) {
  FOSVOS_REQUIRE(workspace && dw && Cin > 0 && Cin <= CinP && Cout > 0 && Cout <= CoutP, "conv3x3_wgrad_tc_finish: bad arguments");
  cudaStream_t st = as_stream(stream);
  wgrad_unpack_kernel<<<min(num_sms() * 8, ceil_div(Cout * Cin, 256)), 256, 0, st>>>((float*)workspace, dw, Cout, Cin, CoutP, CinP,
                                                                                    wg_x_is_a(CinP, CoutP), zero_workspace);
  return check_launch("wgrad_unpack");
}

int fosvos_conv3x3_wgrad_tc(const void* x, const void* dz, float* dw, float* db, void* workspace, int N, int H, int W,
                            int CinP, int CoutP, int Cin, int Cout, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(dw && workspace && Cin > 0 && Cin <= CinP, "conv3x3_wgrad_tc: null pointer or bad Cin");
  cudaMemsetAsync(workspace, 0, fosvos_conv3x3_wgrad_tc_workspace_bytes(CinP, CoutP), as_stream(stream));
  int rc = fosvos_conv3x3_wgrad_tc_accumulate(x, dz, db, workspace, N, H, W, CinP, CoutP, Cout, stream);
  if (rc) return rc;
  return fosvos_conv3x3_wgrad_tc_finish(workspace, dw, CinP, CoutP, Cin, Cout, 0, stream);
}

}  // extern "C"
