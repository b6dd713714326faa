// 3x3 convolution (stride 1, pad 1) as an implicit GEMM on the 5th-generation tensor cores.
//
//   D[pixel, cout] = sum_{tap, cin} X[pixel + tap, cin] * W[cout, tap, cin]        bf16 x bf16 -> fp32
//
//   GEMM-M = 128 output pixels: a TH x TW spatial patch of one frame (TH*TW = 128)
//   GEMM-N = BN output channels (16..256)
//   GEMM-K = 9 taps x Cin, walked in (tap, 64-channel) slabs
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
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one
// elected thread), warps 2..5 = epilogue (TMEM -> registers -> bias/ReLU/mask/accumulate -> bf16
// NHWC global stores).  The same kernel computes the data gradient (weights packed flipped and
// transposed, epilogue = ReLU mask [+ accumulate]).
#include <stdlib.h>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace fosvos {

constexpr int TC_BM = 128;       // pixels per tile
constexpr int TC_BK = 64;        // channels per K slab (128 B of bf16 = one swizzle row)
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;
// diagnostic flag bits (tools/conv_probe.py): drop one pipeline component to see what bounds a layer
constexpr int DBG_SKIP_A = 1 << 16, DBG_SKIP_B = 1 << 17, DBG_SKIP_STORE = 1 << 18, DBG_SKIP_MMA = 1 << 19;

template <int BN> struct TcCfg {
  static constexpr int B_BYTES = BN * TC_BK * 2;
  static constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;   // power of two for BN in {16,32,64,128,256}
  // BN >= 64: the epilogue stages 64-channel output slabs (128 px x 128 B, SWIZZLE_128B) in two 16 KB buffers for TMA stores
  static constexpr bool STAGED = BN >= 64;
  static constexpr int STAGING_BYTES = STAGED ? 2 * TC_A_BYTES : 0;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*barriers*/ + 1024 /*alignment slack*/;
};

struct TcParams {
  const float* bias;             // CoutP fp32 or null
  const __nv_bfloat16* mask;     // (N,H,W,CoutP) or null
  __nv_bfloat16* y;              // (N,H,W,CoutP)
  int N, H, W, CoutP;
  int tiles_x, tiles_y, n_tiles_n, total_tiles;
  int tw_shift;                  // TW = 1 << tw_shift, TH = 128 >> tw_shift
  int k_chunks;                  // ceil(CinP / 64)
  int cin_pad;                   // k_chunks * 64: per-tap K extent of the packed weight
  int flags;
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_y, const TcParams p) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* staging = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::STAGING_BYTES);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* tmem_full = empty_bar + Cfg::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  // warp index through a shuffle: the compiler then treats the role branches as warp-uniform and keeps the
  // descriptors / barrier addresses of the single-thread issue loops in uniform registers (a `lane == 0` branch
  // costs ~2x in tcgen05.mma issue rate: tools/exp/mma_issue.cu)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_w);
    if (Cfg::STAGED) ptx::prefetch_tensormap(&map_y);
    for (int i = 0; i < Cfg::STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 4);     // one arrive per epilogue warp
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int num_kb = 9 * p.k_chunks;
  const int TW = 1 << p.tw_shift, TH = TC_BM >> p.tw_shift;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====================
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int nt = tile % p.n_tiles_n;
        int m = tile / p.n_tiles_n;
        const int tx = m % p.tiles_x; m /= p.tiles_x;
        const int ty = m % p.tiles_y;
        const int n = m / p.tiles_y;
        const int x0 = tx * TW, y0 = ty * TH, n0 = nt * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          const int tap = kb / p.k_chunks, chunk = kb - tap * p.k_chunks;
          const int r = tap / 3, s = tap - 3 * r;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* a_dst = tiles + stage * Cfg::STAGE_BYTES;
          uint8_t* b_dst = a_dst + TC_A_BYTES;
          if (ptx::elect_one()) {
            const bool la = !(p.flags & DBG_SKIP_A), lb = !(p.flags & DBG_SKIP_B);
            if (la || lb) ptx::mbar_expect_tx(&full_bar[stage], (la ? TC_A_BYTES : 0) + (lb ? Cfg::B_BYTES : 0));
            else ptx::mbar_arrive(&full_bar[stage]);
            if (la) ptx::tma_load_4d(a_dst, &map_x, &full_bar[stage], chunk * TC_BK, x0 + s - 1, y0 + r - 1, n);
            if (lb) ptx::tma_load_2d(b_dst, &map_w, &full_bar[stage], tap * p.cin_pad + chunk * TC_BK, n0);
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform loop, one elected lane issues) =====================
    {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(TC_BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1);      // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BN;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);        // TMA bytes have landed
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(tiles + stage * Cfg::STAGE_BYTES);
          const uint64_t da = ptx::umma_desc_sw128_kmajor(a_addr);
          const uint64_t db = ptx::umma_desc_sw128_kmajor(a_addr + TC_A_BYTES);
          if (ptx::elect_one()) {
            if (!(p.flags & DBG_SKIP_MMA)) {
#pragma unroll
              for (int k = 0; k < TC_BK / 16; ++k) {
                // advance 16 bf16 = 32 B along K inside the swizzle row: +2 in the (addr >> 4) field
                ptx::umma_bf16(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
              }
            }
            ptx::umma_commit(&empty_bar[stage]);          // frees the smem slot when the MMAs retire
            if (kb == num_kb - 1) ptx::umma_commit(&tmem_full[as]);   // accumulator complete -> epilogue
          }
          __syncwarp();
          if (++stage == Cfg::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quad = warp & 3;                            // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;                     // accumulator row = pixel inside the patch
    const int py = row >> p.tw_shift, px = row & (TW - 1);
    const bool issuer = (warp == 2) && (lane == 0);       // owns the TMA-store bulk groups
    int it = 0;
    uint32_t slab_ctr = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int nt = tile % p.n_tiles_n;
      int m = tile / p.n_tiles_n;
      const int tx = m % p.tiles_x; m /= p.tiles_x;
      const int ty = m % p.tiles_y;
      const int n = m / p.tiles_y;
      const int x0 = tx * TW, y0 = ty * TH;
      const int gx = x0 + px, gy = y0 + py, n0 = nt * BN;
      const bool in_img = gx < p.W && gy < p.H;
      const long long pix = ((long long)n * p.H + gy) * p.W + gx;
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * BN;
      if constexpr (Cfg::STAGED) {
        // 64-channel slabs: TMEM -> registers -> bias/ReLU/mask/accumulate -> bf16 -> swizzled smem -> one TMA store
        // (the store clips the patch at the frame edge and the channel tail)
        constexpr int SLABS = BN / 64;
#pragma unroll 1
        for (int j = 0; j < SLABS; ++j) {
          const int co0 = n0 + 64 * j;
          const bool live = co0 < p.CoutP;                // uniform: N-tail tiles have dead slabs
          uint32_t packed[32];
          if (live) {
            uint32_t r[64];
            ptx::tmem_ld32(taddr + 64 * j, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
            ptx::tmem_ld32(taddr + 64 * j + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            ptx::tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              const int co = co0 + 8 * h;
              float v[8];
#pragma unroll
              for (int q = 0; q < 8; ++q) v[q] = __uint_as_float(r[8 * h + q]);
              if (co < p.CoutP) {
                if (p.flags & FOSVOS_CONV_BIAS) {
                  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + co));
                  const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + co + 4));
                  v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
                  v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
                }
                if (p.flags & FOSVOS_CONV_RELU) {
#pragma unroll
                  for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                }
                if ((p.flags & (FOSVOS_CONV_MASK | FOSVOS_CONV_ACCUMULATE)) && in_img) {
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
              }
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
                packed[4 * h + q] = *reinterpret_cast<uint32_t*>(&h2);
              }
            }
          }
          if (j == SLABS - 1) {                           // all TMEM reads of this tile are done: hand the accumulator back
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);
          }
          if (live) {
            uint8_t* buf = staging + (slab_ctr & 1) * TC_A_BYTES;
            if (issuer) ptx::tma_store_wait_read<1>();    // the store that used this buffer two slabs ago has read it
            ptx::named_bar_sync(1, 128);
            uint8_t* rowp = buf + row * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) =
                  make_uint4(packed[4 * c], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
            ptx::fence_proxy_async_smem();
            ptx::named_bar_sync(2, 128);
            if (issuer && !(p.flags & DBG_SKIP_STORE)) {
              ptx::tma_store_4d(&map_y, buf, co0, x0, y0, n);
              ptx::tma_store_commit();
            }
            ++slab_ctr;
          }
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 16) {
          uint32_t r[16];
          ptx::tmem_ld16(taddr + c0, r);
          ptx::tmem_ld_wait();
          if (in_img && !(p.flags & DBG_SKIP_STORE)) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int co = n0 + c0 + 8 * h;
              if (co < p.CoutP) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  v[q] = __uint_as_float(r[8 * h + q]);
                  if (p.flags & FOSVOS_CONV_BIAS) v[q] += __ldg(p.bias + co + q);
                  if (p.flags & FOSVOS_CONV_RELU) v[q] = fmaxf(v[q], 0.f);
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
                store8(p.y + o, v);
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
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
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

static int encode_act_map(CUtensorMap* m, const void* x, int N, int H, int W, int C, int TW, int TH) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(activations %dx%dx%dx%d box %dx%d) failed: %d", N, H, W, C, TH, TW, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

static int encode_w_map(CUtensorMap* m, const void* w, int rows, int kdim, int BN) {
  EncodeTiledFn enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  cuuint64_t dims[2] = {(cuuint64_t)kdim, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)kdim * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)BN};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(weights %dx%d box %d) failed: %d", rows, kdim, BN, (int)r); return FOSVOS_ERR_DRIVER; }
  return FOSVOS_OK;
}

// pick the 128-pixel patch shape that wastes the fewest out-of-frame pixels
static int pick_tw_shift(int H, int W) {
  int best = 4;
  long long best_area = -1;
  for (int sh = 2; sh <= 6; ++sh) {            // TW = 4..64, TH = 32..2
    const int TW = 1 << sh, TH = TC_BM >> sh;
    const long long area = (long long)ceil_div(H, TH) * TH * ceil_div(W, TW) * TW;
    if (best_area < 0 || area < best_area || (area == best_area && sh == 4)) { best_area = area; best = sh; }
  }
  return best;
}

template <int BN>
static int launch_tc(const CUtensorMap& mx, const CUtensorMap& mw, const CUtensorMap& my, const TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(smem=%d): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
    attr_set = true;
  }
  int grid = min(p.total_tiles, num_sms());
  if (const char* e = getenv("FOSVOS_TC_GRID")) { const int v = atoi(e); if (v > 0) grid = min(grid, v); }
  conv3x3_tc_kernel<BN><<<grid, TC_THREADS, Cfg::SMEM_BYTES, st>>>(mx, mw, my, p);
  return check_launch("conv3x3_tc");
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_conv3x3_tc(const void* x, const void* w_packed, const float* bias, const void* mask, void* y, int N, int H,
                      int W, int Cin, int Cout, int flags, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && w_packed && y && N > 0 && H > 0 && W > 0, "conv3x3_tc: null pointer or empty shape");
  FOSVOS_REQUIRE(Cin % 8 == 0 && Cout % 8 == 0 && Cin > 0 && Cout > 0,
                 "conv3x3_tc: Cin=%d and Cout=%d must be positive multiples of 8 (pad the NHWC tensors)", Cin, Cout);
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_BIAS) || bias, "conv3x3_tc: BIAS flag without bias pointer");
  FOSVOS_REQUIRE(!(flags & FOSVOS_CONV_MASK) || mask, "conv3x3_tc: MASK flag without mask pointer");
  FOSVOS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0,
                 "conv3x3_tc: pointers must be 16-byte aligned");
  TcParams p;
  p.bias = bias;
  p.mask = (const __nv_bfloat16*)mask;
  p.y = (__nv_bfloat16*)y;
  p.N = N; p.H = H; p.W = W; p.CoutP = Cout;
  p.tw_shift = pick_tw_shift(H, W);
  const int TW = 1 << p.tw_shift, TH = TC_BM >> p.tw_shift;
  p.tiles_x = ceil_div(W, TW);
  p.tiles_y = ceil_div(H, TH);
  p.k_chunks = ceil_div(Cin, TC_BK);
  p.cin_pad = p.k_chunks * TC_BK;
  p.flags = flags;
  int BN = 16;
  while (BN < Cout && BN < 256) BN *= 2;
  // small grids: prefer narrower N tiles so that more SMs get work
  const long long m_tiles = (long long)N * p.tiles_x * p.tiles_y;
  while (BN > 64 && m_tiles * ceil_div(Cout, BN) < num_sms()) BN /= 2;
  if (const char* e = getenv("FOSVOS_TC_BN")) { const int v = atoi(e); if (v >= 16 && v <= 256 && (v & (v - 1)) == 0) BN = v; }
  p.n_tiles_n = ceil_div(Cout, BN);
  FOSVOS_REQUIRE(m_tiles * p.n_tiles_n < (1LL << 31), "conv3x3_tc: too many tiles");
  p.total_tiles = (int)(m_tiles * p.n_tiles_n);

  CUtensorMap mx, mw, my;
  int rc = encode_act_map(&mx, x, N, H, W, Cin, TW, TH);
  if (rc) return rc;
  rc = encode_act_map(&my, y, N, H, W, Cout, TW, TH);      // output slabs leave through TMA stores (BN >= 64)
  if (rc) return rc;
  rc = encode_w_map(&mw, w_packed, Cout, 9 * p.cin_pad, BN);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  switch (BN) {
    case 16: return launch_tc<16>(mx, mw, my, p, st);
    case 32: return launch_tc<32>(mx, mw, my, p, st);
    case 64: return launch_tc<64>(mx, mw, my, p, st);
    case 128: return launch_tc<128>(mx, mw, my, p, st);
    default: return launch_tc<256>(mx, mw, my, p, st);
  }
}

}  // extern "C"
