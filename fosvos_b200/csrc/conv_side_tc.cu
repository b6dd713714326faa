// side_prep convolution (3x3, Cout = 16, bias, no ReLU -- reference osvos_vgg.py:42,69) on the tensor cores, with the
// three taps of a kernel ROW stacked along GEMM-N, and optionally the two 1x1 heads of the side chain (score_dsn,
// osvos_vgg.py:75, and this stage's 16 columns of fuse, :81) applied in the epilogue.
//
// Why not the generic kernel (conv_tc.cu, N = 16): one M = 128 tcgen05.mma costs max(N / 2, 32 + N / 4) cycles -- reading the
// 128-row A tile alone takes 32 (tools/exp/mma_side.cu; ~57 per instruction in the round-1 kernel) -- so with N = 16 every one
// of the 9 * Cin / 16 instructions of a tile pays that floor for 8 cycles of tensor work (12 % tensor pipe, ncu r01i).  Here
//
//   D[q, (s, co)] = sum_r sum_c X[q + (r - 1, 0), c] * W[co, r, s, c]           one GEMM, N = 3 * 16 = 48, K = 3 * Cin
//   out[y, x, co] = b[co] + D[(y, x - 1), (0, co)] + D[(y, x), (1, co)] + D[(y, x + 1), (2, co)]
//
// i.e. the horizontal taps become output columns of the SAME instruction (3x fewer instructions, each still at the
// floor), and the three partial sums of an output pixel sit in neighbouring accumulator rows = neighbouring LANES of one
// warp: the epilogue combines them with two warp shuffles per channel -- no shared-memory staging, whose traffic would
// compete with the tensor core's operand reads (a 9-tap stack, N = 144, needs ~120 KB of staging traffic per tile).
//
// Tile = 32 x 4 pixel patch whose first and last columns are halo (30 x 4 outputs, 94 % useful rows).  A operand: ONE
// halo box {64 ch, 32, 4 + 2} per 64-channel slab (TMA, SWIZZLE_128B; out-of-frame pixels are zero-filled = the conv
// padding); vertical tap r is the same shared-memory tile read from r * 32 rows further (4096 B: a multiple of the
// swizzle atom), as in conv_tc.cu's MODE_HALO.  B operand: ALL weights of the layer ([slab][r][(s, co)][64 ch], at most
// 144 KB for Cin = 512) are loaded ONCE per persistent CTA and stay resident in shared memory.
// Two TMEM accumulators (64 columns apart); warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 / 6-9 = two epilogue
// groups draining alternate tiles.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace fosvos {

constexpr int SD_THREADS = 64 + 8 * 32;
constexpr int SD_TW = 32, SD_TH = 4;
constexpr int SD_OUT_W = SD_TW - 2;                       // output columns per tile
constexpr int SD_A_BYTES = (SD_TH + 2) * SD_TW * 128;     // 24576: halo box of one 64-channel slab
constexpr int SD_WT_BYTES = 48 * 128;                     // 6144: weights of one (slab, kernel row): 48 rows (s, co) x 64 ch
constexpr int SD_ACC_STRIDE = 64;                         // TMEM columns between the two accumulators
constexpr int SD_TMEM_COLS = 128;
constexpr int SD_MISC_BYTES = 1024;                       // bias[16], heads[36], barriers, tmem pointer

struct SdParams {
  const float* bias;             // 16 fp32 or null
  const float* heads;            // 36 fp32: score.w[16], score.b, 3 unused, fuse.w[16 i .. 16 i + 15]; or null
  __nv_bfloat16* y;              // (N, H, W, 16) or null
  float2* zs;                    // (N, H, W) {fuse head, score head} or null
  int N, H, W;
  int tiles_x, tiles_y, total_tiles;
  int k_chunks;                  // 64-channel slabs
  int cin_pad;                   // k_chunks * 64: per-tap K extent of the packed weight
  int stages;                    // depth of the halo-box ring
  uint32_t dx_mul, dx_shift, dy_mul, dy_shift;
};

__global__ void __launch_bounds__(SD_THREADS, 1)
conv3x3_side_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, const SdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                              // stages x SD_A_BYTES
  uint8_t* wres = smem + p.stages * SD_A_BYTES;                      // k_chunks x 3 x SD_WT_BYTES
  uint8_t* misc = wres + p.k_chunks * 3 * SD_WT_BYTES;
  float* bias_s = reinterpret_cast<float*>(misc);                    // 16
  float* heads_s = bias_s + 16;                                      // 36 (+ 12 pad)
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
      ptx::mbar_init(&tmem_empty[i], 4);            // one arrive per warp of the epilogue group
    }
    ptx::mbar_init(w_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(tmem_ptr, SD_TMEM_COLS);
  if (threadIdx.x < 16) bias_s[threadIdx.x] = p.bias ? __ldg(p.bias + threadIdx.x) : 0.f;
  if (threadIdx.x >= 32 && threadIdx.x < 32 + 36) heads_s[threadIdx.x - 32] = p.heads ? __ldg(p.heads + threadIdx.x - 32) : 0.f;
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
    // the layer's weights, once: tile (slab kb, kernel row r) = rows (s, co) from the packed [cout][tap][cin_pad] layout,
    // one {64 ch, 16 couts} box per tap
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(w_bar, (uint32_t)(p.k_chunks * 3 * SD_WT_BYTES));
      for (int kb = 0; kb < p.k_chunks; ++kb)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s)
            ptx::tma_load_2d(wres + ((kb * 3 + r) * 3 + s) * 2048, &map_w, w_bar, (3 * r + s) * p.cin_pad + kb * 64, 0);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int m2 = (int)ptx::fast_div((uint32_t)tile, p.dx_mul, p.dx_shift);
      const int tx = tile - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int x0 = tx * SD_OUT_W - 1, y0 = ty * SD_TH;
      for (int kb = 0; kb < p.k_chunks; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        if (ptx::elect_one()) {
          ptx::mbar_expect_tx(&full_bar[stage], SD_A_BYTES);
          ptx::tma_load_4d(ring + stage * SD_A_BYTES, &map_x, &full_bar[stage], kb * 64, x0, y0 - 1, n);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = ptx::umma_idesc_bf16(128, 48);
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
      const uint32_t tmem_d = tmem_base + as * SD_ACC_STRIDE;
      for (int kb = 0; kb < p.k_chunks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_lo = a_lo0 + stage * (SD_A_BYTES >> 4);
          const uint32_t w_lo = w_lo0 + kb * (3 * SD_WT_BYTES >> 4);
#pragma unroll
          for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_lohi(tmem_d, a_lo + r * ((SD_TW * 128) >> 4) + 2 * k, w_lo + r * (SD_WT_BYTES >> 4) + 2 * k, desc_hi, idesc,
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
    const int e = warp - 2;
    const int grp = e >> 2;
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may access = patch row
    const int as = grp;
    int it = grp;
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.total_tiles; tile += 2 * gridDim.x, it += 2) {
      const int m2 = (int)ptx::fast_div((uint32_t)tile, p.dx_mul, p.dx_shift);
      const int tx = tile - m2 * p.tiles_x;
      const int n = (int)ptx::fast_div((uint32_t)m2, p.dy_mul, p.dy_shift);
      const int ty = m2 - n * p.tiles_y;
      const int gx = tx * SD_OUT_W - 1 + lane, gy = ty * SD_TH + quad;
      ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * SD_ACC_STRIDE;
      uint32_t d0[16], d1[16], d2[16];
      ptx::tmem_ld16(taddr, d0);
      ptx::tmem_ld16(taddr + 16, d1);
      ptx::tmem_ld16(taddr + 32, d2);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[as]);      // accumulator handed back before the arithmetic
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float left = __shfl_up_sync(0xffffffffu, __uint_as_float(d0[j]), 1);       // kernel column 0: from pixel x - 1
        const float right = __shfl_down_sync(0xffffffffu, __uint_as_float(d2[j]), 1);    // kernel column 2: from pixel x + 1
        v[j] = (left + __uint_as_float(d1[j])) + (right + bias_s[j]);
      }
      const bool valid = lane >= 1 && lane <= SD_OUT_W && gx < p.W && gy < p.H;
      const long long pix = ((long long)n * p.H + gy) * p.W + gx;
      if (p.zs) {
        float2 z = make_float2(0.f, 0.f), sc = make_float2(heads_s[16], 0.f);
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          z = ptx::ffma2(make_float2(v[j], v[j + 1]), make_float2(heads_s[20 + j], heads_s[21 + j]), z);
          sc = ptx::ffma2(make_float2(v[j], v[j + 1]), make_float2(heads_s[j], heads_s[j + 1]), sc);
        }
        if (valid) p.zs[pix] = make_float2(z.x + z.y, sc.x + sc.y);
      }
      if (p.y && valid) {
        store8(p.y + pix * 16, *reinterpret_cast<const float(*)[8]>(&v[0]));
        store8(p.y + pix * 16 + 8, *reinterpret_cast<const float(*)[8]>(&v[8]));
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, SD_TMEM_COLS);
  }
}

// ---- host side -------------------------------------------------------------------------------
typedef CUresult (*SdEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static SdEncodeTiledFn sd_get_encode() {
  static SdEncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
    fn = reinterpret_cast<SdEncodeTiledFn>(f);
  else
    cudaGetLastError();
  return fn;
}

static void sd_fast_div(int d, uint32_t& mul, uint32_t& shift) {
  uint32_t l = 0;
  while ((1u << l) < (uint32_t)d) ++l;
  mul = (uint32_t)((((uint64_t)1 << 32) * (((uint64_t)1 << l) - (uint64_t)d)) / (uint64_t)d + 1);
  shift = l;
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_conv3x3_side_tc_supported(int Cin) {
  const int k_chunks = ceil_div(Cin, 64);
  const int w_bytes = k_chunks * 3 * SD_WT_BYTES;
  return (Cin > 0 && Cin % 8 == 0 && (227 * 1024 - 1024 - SD_MISC_BYTES - w_bytes) / SD_A_BYTES >= 2) ? 1 : 0;
}

int fosvos_conv3x3_side_tc(const void* x, const void* w_packed, const float* bias, void* y, void* zs, const float* heads, int N,
                           int H, int W, int Cin, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && w_packed && (y || zs) && N > 0 && H > 0 && W > 0, "conv3x3_side_tc: null pointer or empty shape");
  FOSVOS_REQUIRE(!zs || heads, "conv3x3_side_tc: zs output without the head weights");
  FOSVOS_REQUIRE(fosvos_conv3x3_side_tc_supported(Cin), "conv3x3_side_tc: Cin=%d (multiple of 8, weights must fit in shared memory)", Cin);
  FOSVOS_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)zs & 7) == 0,
                 "conv3x3_side_tc: pointers must be 16-byte aligned (zs: 8)");
  SdEncodeTiledFn enc = sd_get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled not available from the driver"); return FOSVOS_ERR_DRIVER; }
  SdParams p;
  p.bias = bias;
  p.heads = heads;
  p.y = (__nv_bfloat16*)y;
  p.zs = (float2*)zs;
  p.N = N; p.H = H; p.W = W;
  p.tiles_x = ceil_div(W, SD_OUT_W);
  p.tiles_y = ceil_div(H, SD_TH);
  const long long tiles = (long long)N * p.tiles_x * p.tiles_y;
  FOSVOS_REQUIRE(tiles < (1LL << 31), "conv3x3_side_tc: too many tiles");
  p.total_tiles = (int)tiles;
  p.k_chunks = ceil_div(Cin, 64);
  p.cin_pad = p.k_chunks * 64;
  const int w_bytes = p.k_chunks * 3 * SD_WT_BYTES;
  p.stages = min(6, (227 * 1024 - 1024 - SD_MISC_BYTES - w_bytes) / SD_A_BYTES);
  sd_fast_div(p.tiles_x, p.dx_mul, p.dx_shift);
  sd_fast_div(p.tiles_y, p.dy_mul, p.dy_shift);

  CUtensorMap mx, mw;
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {64, SD_TW, SD_TH + 2, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(side_prep activations %dx%dx%dx%d) failed: %d", N, H, W, Cin, (int)r); return FOSVOS_ERR_DRIVER; }
  }
  {
    cuuint64_t dims[2] = {(cuuint64_t)9 * p.cin_pad, 16};
    cuuint64_t strides[1] = {(cuuint64_t)9 * p.cin_pad * 2};
    cuuint32_t box[2] = {64, 16};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w_packed), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(side_prep weights, cin_pad %d) failed: %d", p.cin_pad, (int)r); return FOSVOS_ERR_DRIVER; }
  }
  const int smem_bytes = p.stages * SD_A_BYTES + w_bytes + SD_MISC_BYTES + 1024;
  static unsigned long long attr_set = 0;          // one bit per device: function attributes are per device
  if (first_use_on_device(attr_set)) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_side_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { attr_set = 0; set_error("cudaFuncSetAttribute(side_prep smem): %s", cudaGetErrorString(e)); return FOSVOS_ERR_LAUNCH; }
  }
  const int grid = min(p.total_tiles, num_sms());
  conv3x3_side_tc_kernel<<<grid, SD_THREADS, smem_bytes, as_stream(stream)>>>(mx, mw, p);
  return check_launch("conv3x3_side_tc");
}

}  // extern "C"
