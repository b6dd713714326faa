// 3x3 (and 1x1) convolution, stride 1, as an implicit GEMM on the 5th-generation tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]        bf16 x bf16 -> fp32
//
//   GEMM-M = 128 output pixels: a TH x TW spatial patch of one frame (TH*TW = 128)
//   GEMM-N = BN output channels (16..256)
//   GEMM-K = taps x Cin, walked in (tap, 64-channel) slabs
//
// A operand: the NHWC activation tensor is described to TMA as a 4-D tensor (C, W, H, N); the slab
// for tap (r,s) is the box {64, TW, TH, 1} at (c0, x0+s-1, y0+r-1, n).  Out-of-bounds elements are
// zero-filled by TMA, which is exactly the convolution's zero padding (and the channel tail when
// Cin is not a multiple of 64), and the box lands in shared memory as 128 rows of 128 B in the
// SWIZZLE_128B K-major layout tcgen05.mma consumes.  No im2col buffer ever exists.
// B operand: weights pre-packed [cout][tap][cin_pad] (K-major), box {64, BN}.
// Accumulators live in TMEM (2 stages x BN columns) so the epilogue of tile i overlaps the MMAs of
// tile i+1.  Persistent CTAs, one per SM, static round-robin over tiles.
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..17 =
// two epilogue groups of eight warps (group g drains accumulator stage g, i.e. every other tile; two warps
// share a TMEM lane quadrant and split each 64-column slab):
// TMEM -> registers -> bias/ReLU/mask/accumulate -> bf16 -> swizzled shared-memory slab -> TMA store
// (which also clips the patch at the frame edge).  The role branches are warp-uniform and the single
// issuing lane is picked with elect.sync, so descriptors stay in uniform registers (a `lane == 0` branch
// halves the tcgen05.mma issue rate; tools/exp/mma_issue.cu).
// The same kernel computes the data gradient (weights packed flipped and transposed, epilogue = ReLU
// mask [+ accumulate]).
//
// MODE_HALO (3x3, the default): the nine shifted 128-pixel boxes of a tile overlap almost completely, and
// fetching each of them through L2 makes the wide early layers L2->SM bound (9 x 16 KB per 64 channels and
// tile).  Instead one box {64ch, TW, TH+2} per horizontal shift s is loaded (three per tile and channel slab)
// and the three vertical taps r read it at r * TW rows, a multiple of the 1024-byte swizzle atom when
// TW % 8 == 0, so the same SWIZZLE_128B descriptor applies.  Activation boxes and weight tiles travel in two
// independent mbarrier rings (A: 3-6 halo boxes, B: 4-9 weight tiles), both filled by the producer warp in
// the order the MMA warp consumes them.
//
// MODE_C8 serves the first layer (Cin == 8: the 3-channel frame padded to 8).  A pixel is then 16 bytes = one
// row of an un-swizzled K-major core matrix, and the frame is described to TMA as (8 W, H, N), so ONE box
// {(8+2) px x 8 ch, 16+2 rows} per 16 x 8 patch (18 requests of 160 B) holds all nine taps: tap (r, s) is the
// same tile read from byte offset r * 160 + s * 16, with SBO = 160 (next image row).  A K = 16 step pairs two
// taps (LBO = their byte distance); the 64 x 80 weight (9 taps + one zero block) stays resident in shared
// memory and a tile is five MMAs.
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "ptx.cuh"

namespace fosvos {

// conv_stack_tc.cu
bool conv_stack_tc_supported(int Cin, int Cout);
int conv_stack_tc_launch(const void* x, const void* w_packed, const float* bias, const void* mask, void* y, void* y_pool, void* pool_arg,
                         int N, int H, int W, int Cin, int relu, cudaStream_t stream);

constexpr int TC_BM = 128;       // pixels per tile
constexpr int TC_BK = 64;        // channels per K slab (128 B of bf16 = one swizzle row)
constexpr int TC_EPI_WARPS = 16;   // two groups of eight
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_BIAS_BYTES = 2048;  // all CoutP <= 512 biases, staged once per CTA
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
constexpr int TC_SLAB_BYTES = TC_BM * 128;      // one 64-channel output slab of a tile
constexpr int MODE_GENERIC = 0, MODE_C8 = 1, MODE_HALO = 2, MODE_K16 = 3;
// MODE_K16: MODE_HALO for Cin == 16 (the data gradient of side_prep: dA = conv(dsp (16 ch), W^T)).  The generic path pads the
// 16 channels to a 64-channel K slab, i.e. issues four K = 16 MMAs per tap of which three multiply zeros; here a tap is ONE
// MMA: activation boxes {16 ch, TW, TH + 2} and weight boxes {16, BN} land as 32-byte rows in the SWIZZLE_32B K-major layout
// (8-row atoms of 256 B; a vertical tap is r * TW rows = a multiple of the atom further).
__host__ __device__ constexpr bool tc_halo_like(int mode) { return mode == MODE_HALO || mode == MODE_K16; }
__host__ __device__ constexpr int tc_row_bytes(int mode) { return mode == MODE_K16 ? 32 : 128; }   // bytes of one pixel / weight row in a K slab
constexpr int TC_FLAG_SKIP_B = 1 << 21;          // internal (FOSVOS_TC_SKIP_B=1): timing experiment, weight tiles are not loaded
constexpr int TC_FLAG_SKIP_A = 1 << 22;          // internal (FOSVOS_TC_SKIP_A=1): timing experiment, activation boxes are not loaded
constexpr int TC_FLAG_MASK_IN_REGS = 1 << 20;    // internal (FOSVOS_TC_MASK_REGS=1): apply the ReLU mask in the register phase
constexpr int HALO_A_BYTES = 20 * 1024;          // (TH+2) x TW x 128 B: 18 KB for 16x8 patches, 20 KB for 8x16
constexpr int C8_ROW_BYTES = (8 + 2) * 16;       // one image row of the halo tile: 10 pixels x 8 channels
constexpr int C8_BOX_BYTES = (16 + 2) * C8_ROW_BYTES;   // 2880
constexpr int C8_A_BYTES = 3072;                 // box + zeroed tail (the zero-weight K block reads 16 B past the box)
constexpr int C8_KBLOCKS = 10;                   // 9 taps + 1 zero-weight pad -> 5 MMAs of K = 16
constexpr int C8_W_BYTES = C8_KBLOCKS * 64 * 16; // [k block][64 couts][8 ch]

template <int BN, int MODE> struct TcCfg {
  static constexpr int A_BYTES = MODE == MODE_C8 ? C8_A_BYTES : MODE == MODE_HALO ? HALO_A_BYTES : MODE == MODE_K16 ? HALO_A_BYTES / 4 : TC_A_BYTES;
  static constexpr int B_BYTES = MODE == MODE_C8 ? 0 : BN * tc_row_bytes(MODE);
  // MODE_HALO: two rings -- STAGES halo boxes (A) followed by B_STAGES weight tiles (B); otherwise one ring of A+B
  // generic narrow tiles (side_prep, N = 16/32): two 64-channel K blocks per stage -- the MMAs are issue-bound (~57 cycles
  // each), a barrier round trip per four of them is a measurable tax
  static constexpr int K_GROUP = (MODE == MODE_GENERIC && BN <= 32) ? 2 : 1;
  static constexpr int STAGE_BYTES = tc_halo_like(MODE) ? A_BYTES : K_GROUP * (A_BYTES + B_BYTES);
  static constexpr int STAGES = MODE == MODE_K16 ? 6 : MODE == MODE_HALO ? ((BN >= 256) ? 3 : (BN >= 128) ? 4 : (BN >= 64) ? 4 : 6)
                                : MODE == MODE_C8 ? 8 : (BN >= 256) ? 3 : (BN >= 128) ? 5 : (BN >= 64) ? 7 : 5;
  // narrow tiles: one B stage = the three vertical taps of a kernel column (a N = 64 MMA is issue-bound at ~57 cycles, so
  // a barrier round trip per four MMAs costs a third on top; twelve MMAs per wait bring it under 10 %)
  static constexpr int B_GROUP = (MODE == MODE_K16 || (MODE == MODE_HALO && BN <= 64)) ? 3 : 1;
  static constexpr int B_STAGE_BYTES = B_GROUP * B_BYTES;
  static constexpr int B_STAGES = MODE == MODE_K16 ? 6 : MODE == MODE_HALO ? ((BN >= 256) ? 4 : (BN >= 128) ? 6 : 4) : 0;
  static constexpr int RING_BYTES = STAGES * STAGE_BYTES + B_STAGES * B_STAGE_BYTES;
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // power of two for BN in {16,32,64,128,256}
  // BN >= 64: each epilogue group stages 64-channel output slabs (128 px x 128 B, SWIZZLE_128B) for TMA stores
  static constexpr bool STAGED = BN >= 64;
  // slabs in flight per epilogue group.  The first layer is bound by its 52 MB/frame output write: with one 16 KB slab per
  // group a CTA keeps 32 KB of stores in flight (4.7 MB chip-wide), too little to cover the store latency at HBM rate
  // (Little: ~7 TB/s x ~1.5 us = 11 MB); its tiny operand ring leaves room for four slabs per group.
  static constexpr int STAGING_BUFS = MODE == MODE_C8 ? 4 : 1;
  static constexpr int STAGING_BYTES = STAGED ? 2 * STAGING_BUFS * TC_SLAB_BYTES : 0;
  static constexpr int WRES_BYTES = MODE == MODE_C8 ? C8_W_BYTES : 0;
  static constexpr int SMEM_BYTES = RING_BYTES + STAGING_BYTES + WRES_BYTES + TC_BIAS_BYTES + 1024 /*barriers*/ + 1024 /*alignment slack*/;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

struct TcParams {
  const float* bias;             // CoutP fp32 or null
  const __nv_bfloat16* mask;     // (N,H,W,CoutP) or null
  __nv_bfloat16* y;              // (N,H,W,CoutP)
  __nv_bfloat16* y_pool;         // (N,ceil(H/2),ceil(W/2),CoutP) or null: 2x2/2 ceil-mode max pool of y, fused (ReLU outputs only)
  uint32_t* pool_arg;            // (N,ceil(H/2),ceil(W/2),CoutP/32,2) or null: 2-bit window index of the (first) maximum per pooled
                                 // element, as two bit planes per 32 channels {bit 0, bit 1}; index = 2 dy + dx in scan order
  const __nv_bfloat16* w;        // packed weight (MODE_C8 reads it directly)
  int N, H, W, CoutP;
  int tiles_x, tiles_y, n_tiles_n, total_tiles;
  int tw_shift;                  // TW = 1 << tw_shift, TH = 128 >> tw_shift
  int k_chunks;                  // ceil(CinP / 64)
  int cin_pad;                   // k_chunks * 64: per-tap K extent of the packed weight
  int taps;                      // 9 (3x3, pad 1) or 1 (1x1)
  uint32_t dn_mul, dn_shift, dx_mul, dx_shift, dy_mul, dy_shift;   // fast division by n_tiles_n, tiles_x, tiles_y
  int halo_bytes;                // MODE_HALO: (TH+2) * TW * 128
  int flags;
  // split-operand mode (fp32 through the bf16 tensor cores, `fosvos_conv3x3_tc_split`): the activation tensor holds the
  // bf16 terms of an fp32 map side by side, [x1 | x2 | x3] with seg_len channels each; the GEMM-K walk covers one seg_len
  // segment per kept product x_i * w_j, and segment g reads activation term (seg_map >> 4 g) & 15
  int seg_len;                   // 0 = off
  uint32_t seg_map;
  int split_planes;              // output terms written by the slab epilogue (1 = plain bf16 output)
  int plane_stride;              // channels between two output terms (a multiple of 64)
  float* y_f32;                  // narrow epilogue (BN <= 32): fp32 output instead of bf16 (or null)
  // The tensor core chops (does not round) when it adds into its fp32 accumulator: ~2^-25 relative per MMA, same sign every
  // time, so a K-long chain loses (K / 16) * 2^-25 -- invisible next to a bf16 result, but 1e-5 .. 3e-5 per layer for a
  // split-operand convolution.  Split launches therefore keep FOUR accumulators per tile: the leading products x1 w1 alternate
  // between two by 64-channel slab parity (half the chain each), the first-order products (x2 w1, x1 w2; 2^-8 smaller, so their
  // chopping error is too) take the third, the second-order ones the fourth; the epilogue adds them in fp32, smallest first.
  uint32_t ord_map;              // order (0, 1, 2) of GEMM-K segment g in bits [4g, 4g + 4)
  uint32_t acc_used;             // bit a: accumulator a receives at least one MMA per tile
};

// accumulator (0..3) of 64-channel K slab `ks` in a split launch
__device__ __forceinline__ int tc_acc_id(const TcParams& p, int ks) {
  const int per = p.seg_len >> 6;
  const int g = ks / per;
  const int ord = (int)((p.ord_map >> (4 * g)) & 15u);
  return ord == 0 ? ((ks - g * per) & 1) : ord + 1;
}

// activation channel of GEMM-K position c (a multiple of 64) in split-operand mode
__device__ __forceinline__ int tc_act_chan(const TcParams& p, int c) {
  if (p.seg_len == 0) return c;
  const int g = c / p.seg_len;
  return (int)((p.seg_map >> (4 * g)) & 15u) * p.seg_len + (c - g * p.seg_len);
}
// term t (0, 1, 2) of the bf16 expansion of v: x1 = bf16(v), x2 = bf16(v - x1), x3 = bf16(v - x1 - x2)
__device__ __forceinline__ float tc_split_term(float v, int t) {
  for (int i = 0; i < t; ++i) v -= __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// Un-swizzled K-major operand: 8-row x 16-byte core matrices; `lbo` = byte distance between the two K core
// matrices of one K=16 step, `sbo` = byte distance between consecutive 8-row groups.
__device__ __forceinline__ uint64_t umma_desc_noswizzle_kmajor(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// SPLIT: the split-operand instantiation (fp32 through the bf16 tensor cores): remapped activation channels in the
// producers, bf16-term planes or fp32 out of the epilogue.  A separate instantiation so that the plain bf16 kernels keep
// their register allocation.
template <int BN, int MODE, bool SPLIT = false>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  using Cfg = TcCfg<BN, MODE>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* btiles = smem + Cfg::STAGES * Cfg::STAGE_BYTES;      // MODE_HALO: the weight-tile ring
  uint8_t* staging = smem + Cfg::RING_BYTES;
  uint8_t* wres = staging + Cfg::STAGING_BYTES;
  float* bias_s = reinterpret_cast<float*>(wres + Cfg::WRES_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(wres + Cfg::WRES_BYTES + TC_BIAS_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* bfull_bar = tmem_empty + 2;
  uint64_t* bempty_bar = bfull_bar + Cfg::B_STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bempty_bar + Cfg::B_STAGES);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    if (MODE != MODE_C8) ptx::prefetch_tensormap(&map_w);
    if (Cfg::STAGED) ptx::prefetch_tensormap(&map_y);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], TC_EPI_WARPS / 2);     // one arrive per warp of the epilogue group
    }
    for (int i = 0; i < Cfg::B_STAGES; ++i) {
      ptx::mbar_init(&bfull_bar[i], 1);
      ptx::mbar_init(&bempty_bar[i], 1);
    }
    ptx::fence_barrier_init();
  }
  constexpr int ACC_STRIDE = SPLIT ? 4 * BN : BN;                          // TMEM columns of one accumulator stage
  constexpr int TMEM_COLS = SPLIT ? (8 * BN < 32 ? 32 : 8 * BN) : Cfg::TMEM_COLS;
  static_assert(TMEM_COLS <= 512, "TMEM budget");
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
  // everything above overlaps the tail of the previous kernel under programmatic dependent launch; global
  // memory is only touched after the wait
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  for (int i = threadIdx.x; i < TC_BIAS_BYTES / 4; i += TC_THREADS)
    bias_s[i] = ((p.flags & FOSVOS_CONV_BIAS) && i < p.CoutP) ? __ldg(p.bias + i) : 0.f;
  if constexpr (MODE == MODE_C8) {
    // resident weight image [k block][64 couts][8 ch] from the packed [cout][tap][64] layout; block 9 and dead couts = 0
    for (int i = threadIdx.x; i < C8_KBLOCKS * 64; i += TC_THREADS) {
      const int kb = i >> 6, co = i & 63;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (kb < 9 && co < p.CoutP) v = __ldg(reinterpret_cast<const uint4*>(p.w + ((long long)co * 9 + kb) * 64));
      *reinterpret_cast<uint4*>(wres + i * 16) = v;
    }
    // tails of the activation stages: never written by TMA, read (times zero weights) by the last K block
    for (int i = threadIdx.x; i < Cfg::STAGES * ((C8_A_BYTES - C8_BOX_BYTES) / 16); i += TC_THREADS) {
      const int st = i / ((C8_A_BYTES - C8_BOX_BYTES) / 16), j = i % ((C8_A_BYTES - C8_BOX_BYTES) / 16);
      *reinterpret_cast<uint4*>(tiles + st * Cfg::STAGE_BYTES + C8_BOX_BYTES + j * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async_smem();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int TW = 1 << p.tw_shift, TH = TC_BM >> p.tw_shift;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t tiles_a = ptx::smem_u32(tiles), full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    uint32_t a_dst = tiles_a, bar_full = full_a, bar_empty = empty_a;
    // weight-tile ring (MODE_HALO)
    int bstage = 0;
    uint32_t bphase = 0;
    const uint32_t btiles_a = ptx::smem_u32(btiles), bfull_a = ptx::smem_u32(bfull_bar), bempty_a = ptx::smem_u32(bempty_bar);
    uint32_t b_dst = btiles_a, bbar_full = bfull_a, bbar_empty = bempty_a;
    (void)bstage; (void)bphase; (void)b_dst; (void)bbar_full; (void)bbar_empty;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      // tile -> (n-tile, patch x, patch y, frame) with multiply-shift divisions (three real divisions per tile and
      // thread were ~45 % of the epilogue's instructions)
      int m = (int)ptx::fast_div((uint32_t)tile, p.dn_mul, p.dn_shift);
      const int nt = tile - m * p.n_tiles_n;
      int m2 = (int)ptx::fast_div((uint32_t)m, p.dx_mul, p.dx_shift);
      const int tx = m - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH, n0 = nt * BN;
      if constexpr (MODE == MODE_C8) {
        ptx::mbar_wait_a(bar_empty, phase ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx_a(bar_full, C8_BOX_BYTES);
          ptx::tma_load_3d_a(a_dst, &map_x, bar_full, (x0 - 1) * 8, y0 - 1, n);
        }
        __syncwarp();
        a_dst += Cfg::STAGE_BYTES; bar_full += 8; bar_empty += 8;
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; a_dst = tiles_a; bar_full = full_a; bar_empty = empty_a; }
      } else if constexpr (tc_halo_like(MODE)) {
        for (int c = 0; c < p.cin_pad; c += TC_BK) {
          for (int s = 0; s < 3; ++s) {
            ptx::mbar_wait_a(bar_empty, phase ^ 1);
            if (ptx::elect_one()) {
              if (p.flags & TC_FLAG_SKIP_A) {
                ptx::mbar_expect_tx_a(bar_full, 0);
              } else {
                ptx::mbar_expect_tx_a(bar_full, p.halo_bytes);
                ptx::tma_load_4d_a(a_dst, &map_x, bar_full, SPLIT ? tc_act_chan(p, c) : c, x0 + s - 1, y0 - 1, n);
              }
            }
            __syncwarp();
            a_dst += Cfg::STAGE_BYTES; bar_full += 8; bar_empty += 8;
            if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; a_dst = tiles_a; bar_full = full_a; bar_empty = empty_a; }
            for (int r = 0; r < 3; r += Cfg::B_GROUP) {
              ptx::mbar_wait_a(bbar_empty, bphase ^ 1);
              if (ptx::elect_one()) {
                if (p.flags & TC_FLAG_SKIP_B) {
                  ptx::mbar_expect_tx_a(bbar_full, 0);
                } else {
                  ptx::mbar_expect_tx_a(bbar_full, Cfg::B_STAGE_BYTES);
#pragma unroll
                  for (int g = 0; g < Cfg::B_GROUP; ++g)
                    ptx::tma_load_2d_a(b_dst + g * Cfg::B_BYTES, &map_w, bbar_full, (3 * (r + g) + s) * p.cin_pad + c, n0);
                }
              }
              __syncwarp();
              b_dst += Cfg::B_STAGE_BYTES; bbar_full += 8; bbar_empty += 8;
              if (++bstage == Cfg::B_STAGES) { bstage = 0; bphase ^= 1; b_dst = btiles_a; bbar_full = bfull_a; bbar_empty = bempty_a; }
            }
          }
        }
      } else {
        // taps outer, 64-channel chunks inner; no divisions and only running shared-memory addresses in the loop
        const int n_r = p.taps == 9 ? 3 : 1;
        int wk = 0;                                       // K coordinate in the packed weight: tap * cin_pad + chunk * 64
        for (int r = 0; r < n_r; ++r) {
          const int yy = p.taps == 9 ? y0 + r - 1 : y0;
          for (int s = 0; s < n_r; ++s) {
            const int xx = p.taps == 9 ? x0 + s - 1 : x0;
            for (int c = 0; c < p.cin_pad; c += Cfg::K_GROUP * TC_BK) {
              const int ng = min(Cfg::K_GROUP, (p.cin_pad - c) / TC_BK);          // K blocks in this stage (the last may be short)
              ptx::mbar_wait_a(bar_empty, phase ^ 1);
              if (ptx::elect_one()) {
                ptx::mbar_expect_tx_a(bar_full, ng * (Cfg::A_BYTES + Cfg::B_BYTES));
#pragma unroll
                for (int g = 0; g < Cfg::K_GROUP; ++g) {
                  if (g < ng) {
                    ptx::tma_load_4d_a(a_dst + g * (Cfg::A_BYTES + Cfg::B_BYTES), &map_x, bar_full, SPLIT ? tc_act_chan(p, c + g * TC_BK) : c + g * TC_BK, xx, yy, n);
                    ptx::tma_load_2d_a(a_dst + g * (Cfg::A_BYTES + Cfg::B_BYTES) + Cfg::A_BYTES, &map_w, bar_full, wk + g * TC_BK, n0);
                  }
                }
              }
              __syncwarp();
              wk += ng * TC_BK;
              a_dst += Cfg::STAGE_BYTES; bar_full += 8; bar_empty += 8;
              if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; a_dst = tiles_a; bar_full = full_a; bar_empty = empty_a; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(TC_BM, BN);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    // SWIZZLE_128B K-major descriptor: low word = start address >> 4 | LBO(1) << 16, high word constant
    const uint64_t desc0 = MODE == MODE_K16 ? ptx::umma_desc_sw32_kmajor(ptx::smem_u32(tiles)) : ptx::umma_desc_sw128_kmajor(ptx::smem_u32(tiles));
    const uint32_t a_lo0 = (uint32_t)desc0, desc_hi = (uint32_t)(desc0 >> 32);
    const uint32_t full_a = ptx::smem_u32(full_bar), empty_a = ptx::smem_u32(empty_bar);
    uint32_t a_lo = a_lo0, bar_full = full_a, bar_empty = empty_a;
    // weight-tile ring (MODE_HALO)
    int bstage = 0;
    uint32_t bphase = 0;
    const uint32_t b_lo0 = (uint32_t)(MODE == MODE_K16 ? ptx::umma_desc_sw32_kmajor(ptx::smem_u32(btiles)) : ptx::umma_desc_sw128_kmajor(ptx::smem_u32(btiles)));
    const uint32_t bfull_a = ptx::smem_u32(bfull_bar), bempty_a = ptx::smem_u32(bempty_bar);
    uint32_t b_lo = b_lo0, bbar_full = bfull_a, bbar_empty = bempty_a;
    (void)bstage; (void)bphase; (void)b_lo; (void)bbar_full; (void)bbar_empty;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);      // the epilogue group has drained this accumulator
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + as * ACC_STRIDE;
      uint32_t started = 0;                                 // SPLIT: accumulators that already hold a partial sum of this tile
      (void)started;
      if constexpr (MODE == MODE_C8) {
        ptx::mbar_wait_a(bar_full, phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = ptx::smem_u32(tiles) + stage * Cfg::STAGE_BYTES;
        const uint32_t w_addr = ptx::smem_u32(wres);
        if (ptx::elect_one()) {
#pragma unroll
          for (int j = 0; j < C8_KBLOCKS / 2; ++j) {
            // K blocks 2j, 2j+1 = taps (r, s) = (2j / 3, 2j % 3) and the next one (block 9: zero weights, one pixel on)
            const int r0 = (2 * j) / 3, s0 = (2 * j) % 3;
            const uint32_t lbo = s0 == 2 ? C8_ROW_BYTES - 32 : 16;
            const uint64_t da = umma_desc_noswizzle_kmajor(a_addr + r0 * C8_ROW_BYTES + s0 * 16, lbo, C8_ROW_BYTES);
            const uint64_t db = umma_desc_noswizzle_kmajor(w_addr + 2 * j * 64 * 16, 64 * 16, 128);
            ptx::umma_bf16(tmem_d, da, db, idesc, j != 0);
          }
          ptx::umma_commit_a(bar_empty);
          ptx::umma_commit(&tmem_full[as]);
        }
        __syncwarp();
        bar_full += 8; bar_empty += 8;
        if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; bar_full = full_a; bar_empty = empty_a; }
      } else if constexpr (tc_halo_like(MODE)) {
        const uint32_t tfull = ptx::smem_u32(&tmem_full[as]);
        const uint32_t tap_step = (uint32_t)(TW * tc_row_bytes(MODE)) >> 4;      // one image row of the patch, in descriptor units
        constexpr int K_STEPS = MODE == MODE_K16 ? 1 : TC_BK / 16;                // K = 16 MMAs per slab
        const int n_boxes = 3 * p.k_chunks;
        uint32_t acc = 0;
        for (int bx = 0; bx < n_boxes; ++bx) {
          ptx::mbar_wait_a(bar_full, phase);                // halo box has landed
          uint32_t tmem_t = tmem_d;                         // SPLIT: the accumulator of this K slab (boxes walk slab-major: bx / 3)
          if constexpr (SPLIT) {
            const int a_id = tc_acc_id(p, bx / 3);
            tmem_t = tmem_d + a_id * BN;
            acc = (started >> a_id) & 1u;
            started |= 1u << a_id;
          }
          for (int r = 0; r < 3; r += Cfg::B_GROUP) {
            ptx::mbar_wait_a(bbar_full, bphase);            // weight tile(s) of tap(s) (r.., s) have landed
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
#pragma unroll
              for (int g = 0; g < Cfg::B_GROUP; ++g) {
#pragma unroll
                for (int k = 0; k < K_STEPS; ++k) {
                  ptx::umma_bf16_lohi(tmem_t, a_lo + (r + g) * tap_step + 2 * k, b_lo + g * (Cfg::B_BYTES >> 4) + 2 * k, desc_hi, idesc,
                                      acc | g | k);
                }
              }
              ptx::umma_commit_a(bbar_empty);
              if (r + Cfg::B_GROUP >= 3) {
                ptx::umma_commit_a(bar_empty);
                if (bx == n_boxes - 1) ptx::umma_commit_a(tfull);
              }
            }
            __syncwarp();
            acc = 1;
            b_lo += Cfg::B_STAGE_BYTES >> 4; bbar_full += 8; bbar_empty += 8;
            if (++bstage == Cfg::B_STAGES) { bstage = 0; bphase ^= 1; b_lo = b_lo0; bbar_full = bfull_a; bbar_empty = bempty_a; }
          }
          a_lo += Cfg::STAGE_BYTES >> 4; bar_full += 8; bar_empty += 8;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; a_lo = a_lo0; bar_full = full_a; bar_empty = empty_a; }
        }
      } else {
        const uint32_t tfull = ptx::smem_u32(&tmem_full[as]);
        const int stages_per_tap = (p.k_chunks + Cfg::K_GROUP - 1) / Cfg::K_GROUP;
        const int n_st = p.taps * stages_per_tap;
        for (int kb = 0, in_tap = 0; kb < n_st; ++kb) {
          ptx::mbar_wait_a(bar_full, phase);              // TMA bytes have landed
          ptx::tc_fence_after();
          const int ng = min(Cfg::K_GROUP, p.k_chunks - in_tap * Cfg::K_GROUP);
          if (ptx::elect_one()) {
#pragma unroll
            for (int g = 0; g < Cfg::K_GROUP; ++g) {
              if (g < ng) {
                const uint32_t a_g = a_lo + g * ((Cfg::A_BYTES + Cfg::B_BYTES) >> 4);
                uint32_t tmem_t = tmem_d, first = (uint32_t)(kb | g);
                if constexpr (SPLIT) {                    // the accumulator of K slab in_tap * K_GROUP + g
                  const int a_id = tc_acc_id(p, in_tap * Cfg::K_GROUP + g);
                  tmem_t = tmem_d + a_id * BN;
                  first = (started >> a_id) & 1u;
                  started |= 1u << a_id;
                }
#pragma unroll
                for (int k = 0; k < TC_BK / 16; ++k) {
                  // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
                  ptx::umma_bf16_lohi(tmem_t, a_g + 2 * k, a_g + (Cfg::A_BYTES >> 4) + 2 * k, desc_hi, idesc, (first | k) != 0);
                }
              }
            }
            ptx::umma_commit_a(bar_empty);                // frees the smem slot when the MMAs retire
            if (kb == n_st - 1) ptx::umma_commit_a(tfull);     // accumulator complete -> epilogue
          }
          if (++in_tap == stages_per_tap) in_tap = 0;
          __syncwarp();
          a_lo += Cfg::STAGE_BYTES >> 4; bar_full += 8; bar_empty += 8;
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; a_lo = a_lo0; bar_full = full_a; bar_empty = empty_a; }
        }
      }
    }
  } else {
    // ===================== epilogue: group g = warps 2+8g .. 9+8g drains accumulator stage g =====================
    // Two warps share each TMEM lane quadrant of a group and split every 64-column slab in halves: the drain is
    // a long dependent instruction stream per warp, so its throughput scales with the number of warps.
    const int e = warp - 2;
    const int grp = e >> 3;
    const int half = (e >> 2) & 1;                        // which 32 columns of a slab
    const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                     // accumulator row = pixel inside the patch
    const int py = row >> p.tw_shift, px = row & (TW - 1);
    const bool issuer = ((e & 7) == 0) && (lane == 0);    // owns this group's TMA-store bulk groups
    const int bar_a = 1 + 2 * grp, bar_b = 2 + 2 * grp;
    constexpr int GRP_THREADS = 32 * TC_EPI_WARPS / 2;
    uint8_t* const buf0 = staging + grp * Cfg::STAGING_BUFS * TC_SLAB_BYTES;
    int cur_buf = 0;                                       // staging slab of the next store (group-uniform)
    const int as = grp;
    const bool relu = (p.flags & FOSVOS_CONV_RELU) != 0;
    const uint32_t bias_a = ptx::smem_u32(bias_s);
    constexpr int MASK_PER_THREAD = TC_BM / (GRP_THREADS / 8);       // 16-byte mask chunks per thread and slab
    const int tg = threadIdx.x - 64 - grp * GRP_THREADS;            // thread index inside the epilogue group
    uint4 mk[MASK_PER_THREAD];
    // MASK alone (the data gradient) is applied to the staged slab with coalesced loads, see below; the register
    // path handles it only together with ACCUMULATE (one rounding of mask(acc) + old)
    const bool post = (p.flags & (FOSVOS_CONV_ACCUMULATE | TC_FLAG_MASK_IN_REGS)) != 0;
    const bool mask_slab = Cfg::STAGED && (p.flags & FOSVOS_CONV_MASK) && !post;
    int it = grp;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, it += 2) {
      // tile -> (n-tile, patch x, patch y, frame) with multiply-shift divisions (three real divisions per tile and
      // thread were ~45 % of the epilogue's instructions)
      int m = (int)ptx::fast_div((uint32_t)tile, p.dn_mul, p.dn_shift);
      const int nt = tile - m * p.n_tiles_n;
      int m2 = (int)ptx::fast_div((uint32_t)m, p.dx_mul, p.dx_shift);
      const int tx = m - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH;
      const int gx = x0 + px, gy = y0 + py, n0 = nt * BN;
      const bool in_img = gx < p.W && gy < p.H;
      const long long pix = ((long long)n * p.H + gy) * p.W + gx;
      const uint32_t aphase = (it >> 1) & 1;
      // ReLU-backward mask (= the layer input's post-ReLU activation; x != 0 <=> x > 0) of the tile's first slab: thread ->
      // (16-byte chunk tg & 7, pixel rows tg >> 3, + 32, ...), so a warp reads four pixels x 128 contiguous bytes.  Issued
      // BEFORE the wait for the accumulator: the loads fly while the tensor cores still work on this tile.
      auto load_mask = [&](int co_slab) {
#pragma unroll
        for (int jj = 0; jj < MASK_PER_THREAD; ++jj) {
          const int rw = (tg >> 3) + jj * (GRP_THREADS / 8);
          const int my = y0 + (rw >> p.tw_shift), mx = x0 + (rw & (TW - 1));
          const int co = co_slab + 8 * (tg & 7);
          mk[jj] = (mx < p.W && my < p.H && co < p.CoutP)
                       ? __ldg(reinterpret_cast<const uint4*>(p.mask + (((long long)n * p.H + my) * p.W + mx) * p.CoutP + co))
                       : make_uint4(~0u, ~0u, ~0u, ~0u);          // clipped by the TMA store anyway
        }
      };
      if (mask_slab) load_mask(n0);
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * ACC_STRIDE;
      if constexpr (Cfg::STAGED) {
        // 64-channel slabs: TMEM -> registers -> bias/ReLU/mask/accumulate -> bf16 -> swizzled smem -> one TMA store
        // (the store clips the patch at the frame edge and the channel tail)
        constexpr int SLABS = BN / 64;
        const int n_planes = SPLIT ? p.split_planes : 1;  // the bf16 terms of the fp32 result, written side by side
#pragma unroll 1
        for (int jp = 0; jp < SLABS * n_planes; ++jp) {
          const int j = jp / n_planes, plane = jp - j * n_planes;
          const int co0 = n0 + 64 * j;
          const bool live = co0 < p.CoutP;                // uniform: N-tail tiles have dead slabs
          uint32_t packed[16];
          if (live) {
            uint32_t r[32];
            if constexpr (SPLIT) {
              // the four partial accumulators of the tile, added in fp32 smallest first (second order, first order, the
              // two halves of the leading products)
              bool have = false;
#pragma unroll
              for (int a = 3; a >= 0; --a) {
                if ((p.acc_used >> a) & 1u) {
                  uint32_t t[32];
                  ptx::tmem_ld32(taddr + a * BN + 64 * j + 32 * half, t);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int q = 0; q < 32; ++q) r[q] = have ? __float_as_uint(__uint_as_float(r[q]) + __uint_as_float(t[q])) : t[q];
                  have = true;
                }
              }
            } else {
              ptx::tmem_ld32(taddr + 64 * j + 32 * half, r);
              ptx::tmem_ld_wait();
            }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int co = co0 + 32 * half + 8 * h;    // < 512: bias_s holds zeros past CoutP
              float v[8];
              const uint4 b0 = ptx::lds128(bias_a + 4 * co), b1 = ptx::lds128(bias_a + 4 * co + 16);
              v[0] = __uint_as_float(r[8 * h]) + __uint_as_float(b0.x); v[1] = __uint_as_float(r[8 * h + 1]) + __uint_as_float(b0.y);
              v[2] = __uint_as_float(r[8 * h + 2]) + __uint_as_float(b0.z); v[3] = __uint_as_float(r[8 * h + 3]) + __uint_as_float(b0.w);
              v[4] = __uint_as_float(r[8 * h + 4]) + __uint_as_float(b1.x); v[5] = __uint_as_float(r[8 * h + 5]) + __uint_as_float(b1.y);
              v[6] = __uint_as_float(r[8 * h + 6]) + __uint_as_float(b1.z); v[7] = __uint_as_float(r[8 * h + 7]) + __uint_as_float(b1.w);
              if (SPLIT) {
                // split output: term `plane` of the (ReLU'd) fp32 value; the accumulator is re-read per term instead of
                // keeping 32 more registers live
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = tc_split_term(relu ? fmaxf(v[q], 0.f) : v[q], plane);
#pragma unroll
                for (int q = 0; q < 4; ++q) packed[4 * h + q] = ptx::cvt_bf16x2(v[2 * q], v[2 * q + 1]);
              } else if (post) {
                if (relu) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                }
                if (in_img && co < p.CoutP) {
                  const long long o = pix * p.CoutP + co;
                  if (p.flags & FOSVOS_CONV_MASK) {
                    float mk[8];
                    load8(p.mask + o, mk);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = mk[q] > 0.f ? v[q] : 0.f;
                  }
                  if (p.flags & FOSVOS_CONV_ACCUMULATE) {
                    float old[8];
                    load8(p.y + o, old);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += old[q];
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) packed[4 * h + q] = ptx::cvt_bf16x2(v[2 * q], v[2 * q + 1]);
              } else if (relu) {
#pragma unroll
                for (int q = 0; q < 4; ++q) packed[4 * h + q] = ptx::cvt_bf16x2_relu(v[2 * q], v[2 * q + 1]);
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) packed[4 * h + q] = ptx::cvt_bf16x2(v[2 * q], v[2 * q + 1]);
              }
            }
            if (p.y_pool) {
              // nn.MaxPool2d(2, 2, ceil_mode=True) on the way out (osvos_vgg.py:90).  A warp holds whole image rows of
              // the patch, so the 2x2 partners are lanes +1 and +TW.  Out-of-frame pixels count as 0: the values are
              // post-ReLU, so a zero never wins against an in-frame value and ceil-mode windows come out right.
              // With `pool_arg` the window index of the maximum is recorded too (what max_pool2d_with_indices keeps: the FIRST
              // maximum in scan order, i.e. a later candidate wins only if strictly greater; non-negative bf16 compare as
              // integers), so the pool's backward pass never re-reads the full-resolution activation.
              uint32_t m[16];
              uint32_t right_wins = 0u, row2_wins = 0u;        // bit 2 i + h: channel 2 i + h of this thread's 32
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const uint32_t mine = in_img ? packed[i] : 0u;
                const uint32_t right = __shfl_down_sync(0xffffffffu, mine, 1);
                __nv_bfloat162 a = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&mine), *reinterpret_cast<const __nv_bfloat162*>(&right));
                const uint32_t am = *reinterpret_cast<uint32_t*>(&a);
                const uint32_t below = __shfl_down_sync(0xffffffffu, am, TW);
                __nv_bfloat162 b = __hmax2(a, *reinterpret_cast<const __nv_bfloat162*>(&below));
                m[i] = *reinterpret_cast<uint32_t*>(&b);
                if (p.pool_arg) {
                  right_wins |= ((right & 0xffffu) > (mine & 0xffffu) ? 1u : 0u) << (2 * i) | ((right >> 16) > (mine >> 16) ? 1u : 0u) << (2 * i + 1);
                  row2_wins |= ((below & 0xffffu) > (am & 0xffffu) ? 1u : 0u) << (2 * i) | ((below >> 16) > (am >> 16) ? 1u : 0u) << (2 * i + 1);
                }
              }
              if (p.pool_arg) {
                const uint32_t right_wins_below = __shfl_down_sync(0xffffffffu, right_wins, TW);
                if (in_img && !(px & 1) && !(py & 1) && co0 + 32 * half < p.CoutP) {
                  const int PH = (p.H + 1) >> 1, PW = (p.W + 1) >> 1;
                  uint32_t* dst = p.pool_arg + ((((long long)n * PH + (gy >> 1)) * PW + (gx >> 1)) * (p.CoutP >> 5) + ((co0 + 32 * half) >> 5)) * 2;
                  *reinterpret_cast<uint2*>(dst) = make_uint2((row2_wins & right_wins_below) | (~row2_wins & right_wins), row2_wins);
                }
              }
              if (in_img && !(px & 1) && !(py & 1)) {
                const int PH = (p.H + 1) >> 1, PW = (p.W + 1) >> 1;
                __nv_bfloat16* dst = p.y_pool + (((long long)n * PH + (gy >> 1)) * PW + (gx >> 1)) * p.CoutP + co0 + 32 * half;
#pragma unroll
                for (int h = 0; h < 4; ++h)
                  if (co0 + 32 * half + 8 * h < p.CoutP)
                    *reinterpret_cast<uint4*>(dst + 8 * h) = make_uint4(m[4 * h], m[4 * h + 1], m[4 * h + 2], m[4 * h + 3]);
              }
            }
          }
          if (jp == SLABS * n_planes - 1) {               // all TMEM reads of this tile are done: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
          }
          if (live && p.y) {                              // (y == nullptr: pool-only launch, nothing but the pooled map leaves)
            uint8_t* const buf = buf0 + cur_buf * TC_SLAB_BYTES;
            const uint32_t buf_a = ptx::smem_u32(buf);
            if (Cfg::STAGING_BUFS > 1) cur_buf = cur_buf + 1 == Cfg::STAGING_BUFS ? 0 : cur_buf + 1;
            // the store that last used this slab has read it (the STAGING_BUFS - 1 younger ones may still be in flight)
            if (issuer) ptx::tma_store_wait_read<Cfg::STAGING_BUFS - 1>();
            ptx::named_bar_sync(bar_a, GRP_THREADS);
            const uint32_t rowp = buf_a + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c)
              ptx::sts128(rowp + (((4 * half + c) ^ (row & 7)) << 4),
                          make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]));
            if (mask_slab) {
              // ReLU backward on the staged slab with the mask chunks fetched ahead (see load_mask)
              ptx::named_bar_sync(bar_a, GRP_THREADS);
#pragma unroll
              for (int jj = 0; jj < MASK_PER_THREAD; ++jj) {
                const int rw = (tg >> 3) + jj * (GRP_THREADS / 8);
                const uint32_t sp = buf_a + rw * 128 + (((tg & 7) ^ (rw & 7)) << 4);
                uint4 v = ptx::lds128(sp);
                v.x &= __vcmpne2(mk[jj].x, 0u); v.y &= __vcmpne2(mk[jj].y, 0u); v.z &= __vcmpne2(mk[jj].z, 0u); v.w &= __vcmpne2(mk[jj].w, 0u);
                ptx::sts128(sp, v);
              }
              if (j + 1 < SLABS) load_mask(n0 + 64 * (j + 1));      // in flight during the next slab's TMEM load and math (never with split output)
            }
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(bar_b, GRP_THREADS);
            if (issuer) {
              ptx::tma_store_4d(&map_y, buf, SPLIT ? co0 + plane * p.plane_stride : co0, x0, y0, n);
              ptx::tma_store_commit();
            }
          }
        }
      } else {
        if (half == 0) {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 16) {
            uint32_t r[16];
            if constexpr (SPLIT) {
              bool have = false;
#pragma unroll
              for (int a = 3; a >= 0; --a) {
                if ((p.acc_used >> a) & 1u) {
                  uint32_t t[16];
                  ptx::tmem_ld16(taddr + a * BN + c0, t);
                  ptx::tmem_ld_wait();
#pragma unroll
                  for (int q = 0; q < 16; ++q) r[q] = have ? __float_as_uint(__uint_as_float(r[q]) + __uint_as_float(t[q])) : t[q];
                  have = true;
                }
              }
            } else {
              ptx::tmem_ld16(taddr + c0, r);
              ptx::tmem_ld_wait();
            }
            if (in_img) {
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const int co = n0 + c0 + 8 * h;
                if (co < p.CoutP) {
                  float v[8];
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    v[q] = __uint_as_float(r[8 * h + q]) + bias_s[co + q];
                    if (relu) v[q] = fmaxf(v[q], 0.f);
                  }
                  const long long o = pix * p.CoutP + co;
                  if (p.flags & FOSVOS_CONV_MASK) {
                    float mk[8];
                    load8(p.mask + o, mk);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = mk[q] > 0.f ? v[q] : 0.f;
                  }
                  if (p.flags & FOSVOS_CONV_ACCUMULATE) {
                    float old[8];
                    load8(p.y + o, old);
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] += old[q];
                  }
                  if (SPLIT && p.y_f32) store8(p.y_f32 + o, v); else store8(p.y + o, v);
                }
              }
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
      }
    }
    if (Cfg::STAGED && issuer) ptx::tma_store_wait_all<0>();   // shared memory must outlive the bulk reads
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
    else
      cudaGetLastError();
  });
  return fn;
}

// (C, W, H, N) view of an NHWC tensor; box = {box_c channels, TW, TH, 1}
static int encode_act_map(CUtensorMap* m, const void* x, int N, int H, int W, int C, int TW, int TH, int box_c, bool swizzle, bool sw32 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw32 ? CU_TENSOR_MAP_SWIZZLE_32B : swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations %dx%dx%dx%d box %dx%dx%d) failed: %d", N, H, W, C, TH, TW, box_c, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

// first layer: the (N,H,W,8) frame as (8 W, H, N); box = one 16 x 8 patch with its halo, 10 pixels x 18 rows
static int encode_c8_map(CUtensorMap* m, const void* x, int N, int H, int W) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[3] = {(cuuint64_t)W * 8, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)W * 16, (cuuint64_t)H * W * 16};
  cuuint32_t box[3] = {80, 18, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(first layer %dx%dx%d) failed: %d", N, H, W, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

static int encode_w_map(CUtensorMap* m, const void* w, int rows, int kdim, int BN, bool k16 = false) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
  cuuint32_t box[2] = {(cuuint32_t)(k16 ? 16 : TC_BK), (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, k16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %dx%d box %d) failed: %d", rows, kdim, BN, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

// pick the 128-pixel patch shape that wastes the fewest out-of-frame pixels
static int pick_tw_shift(int H, int W, int sh_min, int sh_max) {
  int best = 4;
  long long best_area = -1;
  for (int sh = sh_min; sh <= sh_max; ++sh) {  // TW = 4..64, TH = 32..2
    const int TW = 1 << sh, TH = TC_BM >> sh;
    const long long area = (long long)ceil_div(H, TH) * TH * ceil_div(W, TW) * TW;
    if (best_area < 0 || area < best_area || (area == best_area && sh == 4)) { best_area = area; best = sh; }
  }
  return best;
}

// Programmatic dependent launch (the kernel's prologue -- barrier init, TMEM allocation, tensor-map prefetch -- overlaps the
// tail of its predecessor; griddepcontrol.wait precedes the first global access).  Measured on the sequence job: +3 % on the
// eagerly launched inference batches, -2 % on the graph-captured fine-tune window, so by default it is applied to launches
// outside stream capture only.  FOSVOS_PDL=0 / 1 forces it off / on everywhere.
static bool pdl_enabled(cudaStream_t st) {
  static int v = -2;
  if (v == -2) { const char* e = getenv("FOSVOS_PDL"); v = e ? (atoi(e) != 0 ? 1 : 0) : -1; }
  if (v >= 0) return v == 1;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess) { cudaGetLastError(); return false; }
  return cs == cudaStreamCaptureStatusNone;
}

template <int BN, int MODE, bool SPLIT = false>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, MODE>;
  static unsigned long long attr_set = 0;          // one bit per device: function attributes are per device
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<BN, MODE, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { attr_set = 0; set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
  }
  int grid = min(p.total_tiles, num_sms());
  if (const char* e = getenv("FOSVOS_TC_GRID")) { const int v = atoi(e); if (v > 0) grid = min(grid, v); }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled(st) ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<BN, MODE, SPLIT>, mx, mw, my, p);
  if (e != cudaSuccess) { set_error("conv3x3_tc launch: %s", cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
  return check_launch("conv3x3_tc");
}

// split-operand launch (fp32 through the bf16 tensor cores): see TcParams
struct TcSplit {
  int seg_len, terms, n_pairs;
  uint32_t seg_map;
  int planes_out, plane_stride;
  float* y_f32;
};

static int conv_tc_common(const void* x, const void* w_packed, const float* bias, const void* mask, void* y, void* y_pool, int N,
                          int H, int W, int Cin, int Cout, int taps, int flags, fosvos_stream_t stream, const char* what,
                          const TcSplit* split = nullptr, void* pool_arg = nullptr) {
  FOSVOS_REQUIRE(x && w_packed && (y || y_pool || (split && split->y_f32)) && N > 0 && H > 0 && W > 0, "%s: null pointer or empty shape", what);
  FOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cin > 0 && Cout > 0,
                 "%s: Cin=%d and Cout=%d must be positive multiples of 8 (pad the NHWC tensors)", what, Cin, Cout);
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_BIAS) || bias, "%s: BIAS flag without bias pointer", what);
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_MASK) || mask, "%s: MASK flag without mask pointer", what);
  FOSVOS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "%s: pointers must be 16-byte aligned", what);
  // 64 output channels (conv1_2 forward, the data gradients of conv1_2 / conv2_1): the row-stacked kernel of conv_stack_tc.cu
  // (three taps per N = 192 MMA instead of three operand-read-bound N = 64 ones)
  // (its epilogue -- two shuffles per output element -- outlasts the MMAs of a ONE-slab tile when it also reads a mask:
  //  conv1_2's data gradient, 205 us against 186 us at batch 5, stays on the generic kernel; FOSVOS_TC_STACK_ALL overrides)
  if (!split && taps == 9 && conv_stack_tc_supported(Cin, Cout) && !(flags & FOSVOS_CONV_ACCUMULATE) &&
      (!(flags & FOSVOS_CONV_MASK) || Cin >= 128 || getenv("FOSVOS_TC_STACK_ALL")) &&
      !((flags & FOSVOS_CONV_MASK) && (flags & FOSVOS_CONV_RELU)) && (((uintptr_t)mask | (uintptr_t)y | (uintptr_t)y_pool) & 31) == 0 &&      // its epilogue moves 32 bytes per access
      !getenv("FOSVOS_TC_NO_STACK"))
    return conv_stack_tc_launch(x, w_packed, (flags & FOSVOS_CONV_BIAS) ? bias : nullptr, (flags & FOSVOS_CONV_MASK) ? mask : nullptr, y, y_pool,
                                pool_arg, N, H, W, Cin, (flags & FOSVOS_CONV_RELU) ? 1 : 0, as_stream(stream));
  TcParams p;
  p.bias = bias;
  p.mask = (const __nv_bfloat16*)mask;
  p.y = (__nv_bfloat16*)y;
  p.y_pool = (__nv_bfloat16*)y_pool;
  p.pool_arg = (uint32_t*)pool_arg;
  p.w = (const __nv_bfloat16*)w_packed;
  p.N = N; p.H = H; p.W = W; p.CoutP = Cout;
  p.seg_len = split ? split->seg_len : 0;
  p.seg_map = split ? split->seg_map : 0u;
  p.split_planes = split ? split->planes_out : 1;
  p.plane_stride = split ? split->plane_stride : 0;
  p.y_f32 = split ? split->y_f32 : nullptr;
  p.ord_map = 0; p.acc_used = 1;
  if (split) {
    static const int order[6] = {0, 1, 1, 2, 2, 2};
    for (int g = 0; g < split->n_pairs; ++g) p.ord_map |= (uint32_t)order[g] << (4 * g);
    p.acc_used = 1u | (split->seg_len > 64 ? 2u : 0u) | (split->terms >= 2 ? 4u : 0u) | (split->terms >= 3 ? 8u : 0u);
  }
  // in split mode `Cin` is the GEMM-K extent per tap (n_pairs segments); the activation tensor holds `terms` segments
  const int Cact = split ? split->terms * split->seg_len : Cin;
  const int Cy = split ? split->planes_out * split->plane_stride : Cout;
  const bool c8 = !split && taps == 9 && Cin == 8 && Cout <= 64 && !getenv("FOSVOS_TC_NO_C8");
  // narrow outputs (side_prep, N = 16) are bound by the MMA issue rate, not by L2 traffic: they keep one ring
  const bool halo = taps == 9 && !c8 && Cout > 32 && !getenv("FOSVOS_TC_NO_HALO");
  // 16 input channels (the side_prep data gradient): one K = 16 MMA per tap instead of a zero-padded 64-channel slab
  const bool k16 = halo && !split && Cin == 16 && !getenv("FOSVOS_TC_NO_K16");
  p.tw_shift = c8 ? 3 : halo ? pick_tw_shift(H, W, 3, 4) : pick_tw_shift(H, W, 2, 6);
  const int TW = 1 << p.tw_shift, TH = TC_BM >> p.tw_shift;
  p.halo_bytes = (TH + 2) * TW * (k16 ? 32 : 128);
  p.tiles_x = ceil_div(W, TW);
  p.tiles_y = ceil_div(H, TH);
  p.k_chunks = ceil_div(Cin, TC_BK);
  p.cin_pad = p.k_chunks * TC_BK;
  p.taps = taps;
  p.flags = flags & 0xffff;
  if ((flags & FOSVOS_CONV_MASK) && getenv("FOSVOS_TC_MASK_REGS")) p.flags |= TC_FLAG_MASK_IN_REGS;
  if (getenv("FOSVOS_TC_SKIP_B")) p.flags |= TC_FLAG_SKIP_B;
  if (getenv("FOSVOS_TC_SKIP_A")) p.flags |= TC_FLAG_SKIP_A;
  const long long m_tiles = (long long)N * p.tiles_x * p.tiles_y;
  // N tile: minimise waves x cycles per tile.  One M=128 tcgen05.mma costs max(N/2, 32 + N/4) cycles with resident operands
  // (tools/exp/mma_side.cu); in this kernel a narrow one was measured at ~57 (round 1), the constant kept below
  // (tools/exp/mma_issue.cu), so tiles narrower than 128 only pay when they fill an otherwise idle machine.
  int BN = 16;
  {
    int cap = 16;
    while (cap < Cout && cap < 256) cap *= 2;
    double best_cost = -1.0;
    for (int bn = cap; bn >= 64; bn >>= 1) {
      const long long waves = ceil_div_ll(m_tiles * ceil_div(Cout, bn), num_sms());
      const double cost = (double)waves * (bn >= 128 ? bn / 2 : 57);
      // wider tiles re-read less of the activation and drain fewer accumulators: a narrower one must win clearly
      if (best_cost < 0 || cost < 0.92 * best_cost) { best_cost = cost; BN = bn; }
    }
    if (cap < 64) BN = cap;
  }
  if (const char* e = getenv("FOSVOS_TC_BN")) { const int v = atoi(e); if (v >= 16 && v <= 256 && (v & (v - 1)) == 0) BN = v; }
  if (c8) BN = 64;
  if (split && BN > 64) BN = 64;          // four accumulators per tile and two tiles in flight: 8 x 64 = the 512 TMEM columns
  p.n_tiles_n = ceil_div(Cout, BN);
  {
    auto fd = [](int d, uint32_t& mul, uint32_t& shift) {
      uint32_t l = 0;
      while ((1u << l) < (uint32_t)d) ++l;
      mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - (uint64_t)d)) / (uint64_t)d + 1);
      shift = l;
    };
    fd(p.n_tiles_n, p.dn_mul, p.dn_shift);
    fd(p.tiles_x, p.dx_mul, p.dx_shift);
    fd(p.tiles_y, p.dy_mul, p.dy_shift);
  }
  FOSVOS_REQUIRE(m_tiles * p.n_tiles_n < (1LL << 31), "%s: too many tiles", what);
  p.total_tiles = (int)(m_tiles * p.n_tiles_n);

  CUtensorMap mx, mw, my;
  int rc = c8     ? encode_c8_map(&mx, x, N, H, W)
           : k16  ? encode_act_map(&mx, x, N, H, W, Cin, TW, TH + 2, 16, true, true)
           : halo ? encode_act_map(&mx, x, N, H, W, Cact, TW, TH + 2, TC_BK, true)
                  : encode_act_map(&mx, x, N, H, W, Cact, TW, TH, TC_BK, true);
  if (rc) return rc;
  // output slabs leave through TMA stores (BN >= 64); a pool-only launch (y == nullptr) never issues one: its map
  // just has to encode, so it describes the input tensor
  rc = y ? encode_act_map(&my, y, N, H, W, Cy, TW, TH, TC_BK, true) : encode_act_map(&my, x, N, H, W, Cact, TW, TH, TC_BK, true);
  if (rc) return rc;
  rc = encode_w_map(&mw, w_packed, Cout, taps * p.cin_pad, BN, k16);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  if (c8) return launch_tc<64, MODE_C8>(mx, mw, my, p, st);
  if (split) {
    if (halo) return launch_tc<64, MODE_HALO, true>(mx, mw, my, p, st);
    return BN <= 16 ? launch_tc<16, MODE_GENERIC, true>(mx, mw, my, p, st) : launch_tc<32, MODE_GENERIC, true>(mx, mw, my, p, st);
  }
  if (k16) {
    switch (BN) {
      case 64: return launch_tc<64, MODE_K16>(mx, mw, my, p, st);
      case 128: return launch_tc<128, MODE_K16>(mx, mw, my, p, st);
      default: return launch_tc<256, MODE_K16>(mx, mw, my, p, st);
    }
  }
  if (halo) {
    switch (BN) {
      case 16: return launch_tc<16, MODE_HALO>(mx, mw, my, p, st);
      case 32: return launch_tc<32, MODE_HALO>(mx, mw, my, p, st);
      case 64: return launch_tc<64, MODE_HALO>(mx, mw, my, p, st);
      case 128: return launch_tc<128, MODE_HALO>(mx, mw, my, p, st);
      default: return launch_tc<256, MODE_HALO>(mx, mw, my, p, st);
    }
  }
  switch (BN) {
    case 16: return launch_tc<16, MODE_GENERIC>(mx, mw, my, p, st);
    case 32: return launch_tc<32, MODE_GENERIC>(mx, mw, my, p, st);
    case 64: return launch_tc<64, MODE_GENERIC>(mx, mw, my, p, st);
    case 128: return launch_tc<128, MODE_GENERIC>(mx, mw, my, p, st);
    default: return launch_tc<256, MODE_GENERIC>(mx, mw, my, p, st);
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_conv3x3_tc(const void* x, const void* w_packed, const float* bias, const void* mask, void* y, int N, int H,
                      int W, int Cin, int Cout, int flags, fosvos_stream_t stream) {
  return conv_tc_common(x, w_packed, bias, mask, y, nullptr, N, H, W, Cin, Cout, 9, flags, stream, "conv3x3_tc");
}

int fosvos_split_pairs(int terms) { return terms == 1 ? 1 : terms == 2 ? 3 : terms == 3 ? 6 : -1; }

// products kept for `terms` bf16 terms per operand, in the order the GEMM-K segments are walked (and the weights are packed,
// fosvos_pack_conv3x3_weight_split): everything above 2^-8 terms relative to x1 * w1
//   terms 2: x1 w1, x2 w1, x1 w2                      terms 3: + x3 w1, x2 w2, x1 w3
static const int kSplitActTerm[6] = {0, 1, 0, 2, 1, 0};
int fosvos_split_weight_term(int terms, int pair) {
  static const int w_term[6] = {0, 0, 1, 0, 1, 2};
  return (pair >= 0 && pair < fosvos_split_pairs(terms)) ? w_term[pair] : -1;
}

int fosvos_conv3x3_tc_split(const void* x, const void* w_packed, const float* bias, void* y, float* y_f32, int N, int H, int W,
                            int seg_len, int terms, int Cout, int flags, fosvos_stream_t stream) {
  const int n_pairs = fosvos_split_pairs(terms);
  FOSVOS_REQUIRE(n_pairs > 0 && seg_len > 0 && seg_len % 64 == 0, "conv3x3_tc_split: terms=%d (1..3), seg_len=%d (a multiple of 64)", terms, seg_len);
  FOSVOS_REQUIRE(!(flags & (FOSVOS_CONV_MASK | FOSVOS_CONV_ACCUMULATE)), "conv3x3_tc_split: forward epilogues only (BIAS, RELU)");
  FOSVOS_REQUIRE((y != nullptr) != (y_f32 != nullptr), "conv3x3_tc_split: exactly one of y (split bf16 terms) and y_f32");
  FOSVOS_REQUIRE(y_f32 ? Cout <= 32 : Cout > 32, "conv3x3_tc_split: fp32 output serves the narrow layers (Cout <= 32), split output the wide ones");
  FOSVOS_REQUIRE(((uintptr_t)y_f32 & 15) == 0, "conv3x3_tc_split: y_f32 must be 16-byte aligned");
  TcSplit sp;
  sp.seg_len = seg_len; sp.terms = terms; sp.n_pairs = n_pairs;
  sp.seg_map = 0;
  for (int g = 0; g < n_pairs; ++g) sp.seg_map |= (uint32_t)kSplitActTerm[g] << (4 * g);
  sp.planes_out = y ? terms : 1;
  sp.plane_stride = (Cout + 63) / 64 * 64;
  sp.y_f32 = y_f32;
  return conv_tc_common(x, w_packed, bias, nullptr, y, nullptr, N, H, W, n_pairs * seg_len, Cout, 9, flags, stream, "conv3x3_tc_split", &sp);
}

int fosvos_conv3x3_tc_pool(const void* x, const void* w_packed, const float* bias, void* y, void* y_pool, int N, int H, int W,
                           int Cin, int Cout, int flags, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(y_pool, "conv3x3_tc_pool: y_pool is null");
  FOSVOS_REQUIRE((flags & FOSVOS_CONV_RELU) && !(flags & (FOSVOS_CONV_MASK | FOSVOS_CONV_ACCUMULATE)),
                 "conv3x3_tc_pool: the fused pool needs the RELU epilogue (non-negative values) and no mask/accumulate");
  FOSVOS_REQUIRE(Cout >= 64 && ((uintptr_t)y_pool & 15) == 0, "conv3x3_tc_pool: Cout=%d must be >= 64 (slab epilogue), y_pool 16-byte aligned", Cout);
  return conv_tc_common(x, w_packed, bias, nullptr, y, y_pool, N, H, W, Cin, Cout, 9, flags, stream, "conv3x3_tc_pool");
}

int fosvos_conv3x3_tc_pool_arg(const void* x, const void* w_packed, const float* bias, void* y, void* y_pool, void* pool_arg, int N, int H,
                               int W, int Cin, int Cout, int flags, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(y_pool && pool_arg, "conv3x3_tc_pool_arg: y_pool or pool_arg is null");
  FOSVOS_REQUIRE((flags & FOSVOS_CONV_RELU) && !(flags & (FOSVOS_CONV_MASK | FOSVOS_CONV_ACCUMULATE)),
                 "conv3x3_tc_pool_arg: the fused pool needs the RELU epilogue (non-negative values) and no mask/accumulate");
  FOSVOS_REQUIRE(Cout >= 64 && Cout % 32 == 0 && ((uintptr_t)y_pool & 15) == 0 && ((uintptr_t)pool_arg & 7) == 0,
                 "conv3x3_tc_pool_arg: Cout=%d must be a multiple of 32 and >= 64, y_pool 16-byte and pool_arg 8-byte aligned", Cout);
  return conv_tc_common(x, w_packed, bias, nullptr, y, y_pool, N, H, W, Cin, Cout, 9, flags, stream, "conv3x3_tc_pool_arg", nullptr, pool_arg);
}

}  // extern "C"
