/*
 * fosvos_b200.h -- C ABI of libfosvos_sm100.so
 *
 * The drop-in boundary of the B200-native OSVOS VGG-16 hot path.  The reference
 * (klausondrag/FOSVOS) has no FFI of its own: its arithmetic goes through
 * PyTorch library calls.  Each entry point below replaces one such call group
 * on the path named by BASELINE.json; the reference call site it replaces is
 * cited as file:line into /root/reference/src.  INTEGRATION.md shows the
 * ctypes binding the reference's maintainers would add.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - every pointer is a DEVICE pointer unless named host_*
 *   - the caller owns all memory (incl. workspaces); pointers are borrowed for
 *     the duration of the work enqueued on `stream` (a cudaStream_t)
 *   - nothing synchronises; every call is CUDA-graph capturable
 *   - return 0 on success, a negative fosvos_status otherwise; no exceptions
 *     cross the ABI; fosvos_last_error() gives a human-readable reason
 *   - activations inside the path are NHWC ("pixel-major"): element (n,y,x,c)
 *     lives at ((n*H + y)*W + x)*C + c, C a multiple of 8; dtype is
 *     FOSVOS_F32 or FOSVOS_BF16.  Tensors at the module boundary keep the
 *     reference's NCHW fp32 layout (input frame, the five logit maps).
 *   - there is NO CPU fallback: on a machine without an sm_100 device every
 *     compute entry point returns FOSVOS_ERR_NO_DEVICE / FOSVOS_ERR_ARCH.
 */
#ifndef FOSVOS_B200_H_
#define FOSVOS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fosvos_stream_t; /* cudaStream_t */

enum fosvos_status {
  FOSVOS_OK = 0,
  FOSVOS_ERR_BAD_ARG = -1,   /* shape / alignment / enum out of range            */
  FOSVOS_ERR_ARCH = -2,      /* device is not compute capability 10.x             */
  FOSVOS_ERR_NO_DEVICE = -3, /* no CUDA device / driver                            */
  FOSVOS_ERR_LAUNCH = -4,    /* cudaGetLastError() after a launch                  */
  FOSVOS_ERR_DRIVER = -5,    /* tensor-map encode or other driver-API failure      */
  FOSVOS_ERR_UNSUPPORTED = -6
};

enum fosvos_dtype { FOSVOS_F32 = 0, FOSVOS_BF16 = 1 };

/* epilogue flags of the 3x3 convolution kernels */
enum fosvos_conv_flags {
  FOSVOS_CONV_BIAS = 1,       /* add bias[cout]                                           */
  FOSVOS_CONV_RELU = 2,       /* max(.,0): nn.ReLU(inplace=True), osvos_vgg.py:93         */
  FOSVOS_CONV_MASK = 4,       /* multiply by (mask[n,y,x,cout] > 0): ReLU backward        */
  FOSVOS_CONV_ACCUMULATE = 8  /* y += result instead of y = result (gradient fan-in)      */
};

/* packed layouts of a 3x3 conv weight (source is always OIHW fp32, the
 * reference state_dict layout, osvos_vgg.py:92) */
enum fosvos_wlayout {
  FOSVOS_W_SIMT_FWD = 0,   /* [tap][cin][cout]                       (direct kernels)      */
  FOSVOS_W_SIMT_DGRAD = 1, /* [tap'][cout][cin], tap' = 8 - tap      (direct kernels)      */
  FOSVOS_W_TC_FWD = 2,     /* [cout][tap][cin_pad]   K-major bf16    (tcgen05 kernels)     */
  FOSVOS_W_TC_DGRAD = 3    /* [cin][tap'][cout_pad]  K-major bf16    (tcgen05 kernels)     */
};

/* ---- library / device ------------------------------------------------------------ */
int fosvos_abi_version(void);                 /* bumps when any signature changes      */
const char* fosvos_last_error(void);          /* thread-local, never NULL              */
int fosvos_device_check(int device);          /* FOSVOS_OK iff `device` is sm_100      */
int fosvos_num_sms(int device);               /* >0, or a negative status              */

/* ---- frame ingest ------------------------------------------------------------------
 * NCHW fp32 frame (N,C,H,W) -> NHWC activations (N,H,W,Cp), channels >= C zeroed.
 * Replaces the implicit layout of the tensor handed to net.forward
 * (train_online.py:79, util/experiment_helper.py:46). */
int fosvos_nchw_to_nhwc(const float* x_nchw, void* y_nhwc, int N, int C, int H, int W, int Cp,
                        int dtype, fosvos_stream_t stream);
/* NHWC activations -> NCHW fp32 (introspection path: per-conv outputs for hooks,
 * prune.py:96-103). */
int fosvos_nhwc_to_nchw(const void* x_nhwc, float* y_nchw, int N, int C, int H, int W, int Cp,
                        int dtype, fosvos_stream_t stream);

/* ---- weights -----------------------------------------------------------------------
 * Repack one OIHW fp32 3x3 weight (Cout,Cin,3,3) into a kernel layout (derived cache; never
 * serialised -- the state_dict keeps the reference layout, osvos_vgg.py:50-56).
 * CoutP/CinP are the padded channel counts (multiples of 8) of the NHWC activations the
 * kernels see; entries outside the logical weight are zero, so pruned networks with arbitrary
 * widths (prune.py:490-514) run on the same kernels.  The TC layouts additionally pad their
 * contiguous (GEMM-K) channel dimension to a multiple of 64.
 * fosvos_packed_weight_elems returns the element count of the packed buffer.
 * fosvos_pad_bias copies bias (or zeros if NULL: pruned convs have bias=False, prune.py:500)
 * into a CoutP-long fp32 vector. */
long long fosvos_packed_weight_elems(int CoutP, int CinP, int layout);
int fosvos_pack_conv3x3_weight(const float* w_oihw, void* w_packed, int Cout, int Cin, int CoutP,
                               int CinP, int layout, int dtype, fosvos_stream_t stream);
int fosvos_pad_bias(const float* bias, float* bias_padded, int C, int Cp, fosvos_stream_t stream);

/* ---- 3x3 convolution, stride 1, pad 1 (nn.Conv2d, osvos_vgg.py:42,92) --------------
 * y[n,y,x,co] = epilogue( sum_{tap,ci} x[n,y+r-1,x+s-1,ci] * w[co,ci,r,s] )
 * x: (N,H,W,Cin) NHWC, y: (N,H,W,Cout) NHWC, both `dtype`; bias fp32 or NULL;
 * mask: (N,H,W,Cout) NHWC `dtype` or NULL (FOSVOS_CONV_MASK).
 * The same entry points serve the data gradient: pass the *_DGRAD packed
 * weight, Cin<->Cout swapped, flags = MASK (+ACCUMULATE).
 *   _simt : direct fp32-accumulate kernel, w_packed in FOSVOS_W_SIMT_* layout, `dtype` elements
 *   _tc   : tcgen05/TMEM implicit GEMM fed by TMA, bf16 only, FOSVOS_W_TC_* layout */
int fosvos_conv3x3_simt(const void* x, const void* w_packed, const float* bias, const void* mask,
                        void* y, int N, int H, int W, int Cin, int Cout, int flags, int dtype,
                        fosvos_stream_t stream);
int fosvos_conv3x3_tc(const void* x, const void* w_packed, const float* bias, const void* mask,
                      void* y, int N, int H, int W, int Cin, int Cout, int flags,
                      fosvos_stream_t stream);
/* Same, and additionally y_pool (N,ceil(H/2),ceil(W/2),Cout) = nn.MaxPool2d(2,2,ceil_mode=True)(y)
 * (osvos_vgg.py:90) written from the same epilogue.  Needs FOSVOS_CONV_RELU (values >= 0), Cout >= 64,
 * no MASK / ACCUMULATE. */
int fosvos_conv3x3_tc_pool(const void* x, const void* w_packed, const float* bias, void* y,
                           void* y_pool, int N, int H, int W, int Cin, int Cout, int flags,
                           fosvos_stream_t stream);
/* (y may be NULL: only the pooled map is written -- inference never reads the full-resolution
 * output of a stage's last conv when nothing else consumes it, e.g. conv1_2, osvos_vgg.py:63,68.) */
/* Training form: additionally records WHICH element of each 2x2 window the maximum is -- the index
 * autograd's max_pool2d_with_indices keeps for nn.MaxPool2d's backward (osvos_vgg.py:90): the first maximum
 * in (row, col) scan order, 2 dy + dx.  pool_arg: (N,ceil(H/2),ceil(W/2),Cout/32,2) uint32, two bit planes
 * {bit 0, bit 1} per 32 channels (bit c % 32 = channel c).  Cout % 32 == 0.  fosvos_maxpool2x2_bwd_arg consumes it. */
int fosvos_conv3x3_tc_pool_arg(const void* x, const void* w_packed, const float* bias, void* y,
                               void* y_pool, void* pool_arg, int N, int H, int W, int Cin, int Cout,
                               int flags, fosvos_stream_t stream);

/* side_prep: nn.Conv2d(C, 16, 3, padding=1) with bias and NO ReLU (osvos_vgg.py:42,69) on the tensor
 * cores with the three taps of a kernel row stacked along GEMM-N (N = 48; conv_side_tc.cu).
 * w_packed: FOSVOS_W_TC_FWD layout of the (16,Cin,3,3) weight; bias: 16 fp32 or NULL.
 * y  (N,H,W,16) bf16 or NULL: the side_prep map (needed by the backward pass).
 * zs (N,H,W) float2 or NULL: {sum_c fuse.w[16i+c] * y[c], sum_c score_dsn.w[c] * y[c] + score_dsn.b} computed
 *    from the fp32 accumulators -- the two 1x1 heads of the side chain (osvos_vgg.py:75,81) -- with
 *    heads = params + fosvos_side_params_heads_offset(i) of a prepared side-chain parameter block.
 * At least one of y / zs is non-NULL.  fosvos_conv3x3_side_tc_supported(Cin): 1 when the layer's
 * weights fit the kernel's resident shared-memory image (Cin <= 512). */
int fosvos_conv3x3_side_tc_supported(int Cin);
int fosvos_conv3x3_side_tc(const void* x, const void* w_packed, const float* bias, void* y, void* zs,
                           const float* heads, int N, int H, int W, int Cin, fosvos_stream_t stream);

/* ---- fp32 through the bf16 tensor cores (precision 'fp32_tc': the strict-parity mode on tcgen05) ----------------
 * An fp32 value travels as `terms` bf16 terms x1 = bf16(v), x2 = bf16(v - x1), x3 = bf16(v - x1 - x2) (three terms hold all
 * 24 bits); a convolution keeps the fosvos_split_pairs(terms) term products above 2^-8*terms relative to x1*w1 (3 for two
 * terms, 6 for three) and accumulates them in fp32, which reproduces torch's fp32 convolution with TF32 off
 * (osvos_vgg.py:92 as the reference runs it).  Layout: a split activation map is ONE bf16 NHWC tensor
 * (N,H,W, terms*seg_len), the terms side by side with seg_len (a multiple of 64, >= C) channels each.
 *   fosvos_split_nchw: NCHW fp32 frame -> split map (replaces the layout change of fosvos_nchw_to_nhwc).
 *   fosvos_pack_conv3x3_weight_split: OIHW fp32 -> [CoutP][tap][pairs*seg_len] bf16 whose GEMM-K segment g holds weight
 *     term fosvos_split_weight_term(terms, g) (fosvos_packed_weight_split_elems elements).
 *   fosvos_conv3x3_tc_split: 3x3 conv + bias (+ ReLU) of a split map; Cout > 32: y = split map (N,H,W, terms*ceil64(Cout));
 *     Cout <= 32 (side_prep): y_f32 = plain fp32 NHWC (N,H,W,Cout) for the fp32 side chain.  Exactly one of y / y_f32.
 *   fosvos_maxpool2x2_split: nn.MaxPool2d(2,2,ceil_mode=True) (osvos_vgg.py:90) on a split map (exact for three terms). */
int fosvos_split_pairs(int terms);
int fosvos_split_weight_term(int terms, int pair);
int fosvos_split_nchw(const float* x_nchw, void* y_split, int N, int C, int H, int W, int seg_len, int terms,
                      fosvos_stream_t stream);
long long fosvos_packed_weight_split_elems(int CoutP, int seg_len, int terms);
int fosvos_pack_conv3x3_weight_split(const float* w_oihw, void* w_packed, int Cout, int Cin, int CoutP, int seg_len,
                                     int terms, fosvos_stream_t stream);
int fosvos_conv3x3_tc_split(const void* x_split, const void* w_packed, const float* bias, void* y_split, float* y_f32,
                            int N, int H, int W, int seg_len, int terms, int Cout, int flags, fosvos_stream_t stream);
int fosvos_maxpool2x2_split(const void* x_split, void* y_split, int N, int H, int W, int seg_len, int terms,
                            fosvos_stream_t stream);

/* weight + bias gradient of the same convolution (autograd convolution_backward,
 * train_online.py:93):  dw[co,ci,r,s] += sum_p x[p+tap,ci] * dz[p,co];  db[co] += sum_p dz[p,co]
 * x: (N,H,W,CinP), dz: (N,H,W,CoutP) NHWC; dw: OIHW fp32 (Cout,Cin,3,3) -- the parameter's
 * .grad, accumulated across micro-steps; db (Cout) may be NULL. */
int fosvos_conv3x3_wgrad_simt(const void* x, const void* dz, float* dw_oihw, float* db, int N, int H,
                              int W, int CinP, int CoutP, int Cin, int Cout, int dtype,
                              fosvos_stream_t stream);

/* Same gradient on the tensor cores (bf16 operands, fp32 accumulation in TMEM; GEMM-K = pixels,
 * split-K merged by vectorised fp32 reductions).  `workspace`: fosvos_conv3x3_wgrad_tc_workspace_bytes
 * bytes of scratch ([tap][M][N] fp32), overwritten by the call. */
size_t fosvos_conv3x3_wgrad_tc_workspace_bytes(int CinP, int CoutP);
int fosvos_conv3x3_wgrad_tc(const void* x, const void* dz, float* dw_oihw, float* db, void* workspace,
                            int N, int H, int W, int CinP, int CoutP, int Cin, int Cout,
                            fosvos_stream_t stream);
/* The same in two steps, for gradient accumulation over micro-iterations (train_online.py:93-101,
 * `avg_grad_every_n`): `_accumulate` adds this call's gradient to a ZERO-INITIALISED workspace that
 * the caller keeps live (and adds the bias gradient to db, may be NULL); `_finish` adds the
 * workspace to the OIHW fp32 .grad tensor once per optimizer step and (zero_workspace != 0)
 * clears it for the next round. */
int fosvos_conv3x3_wgrad_tc_accumulate(const void* x, const void* dz, float* db, void* workspace,
                                       int N, int H, int W, int CinP, int CoutP, int Cout,
                                       fosvos_stream_t stream);
int fosvos_conv3x3_wgrad_tc_finish(void* workspace, float* dw_oihw, int CinP, int CoutP, int Cin,
                                   int Cout, int zero_workspace, fosvos_stream_t stream);

/* ---- 2x2 / stride-2 max pooling, ceil_mode=True (nn.MaxPool2d, osvos_vgg.py:90) ---- */
int fosvos_maxpool2x2_fwd(const void* x, void* y, int N, int H, int W, int C, int dtype,
                          fosvos_stream_t stream);
/* dx = dy routed to the first maximum of each window in (row, col) scan order, zero
 * elsewhere (what autograd's max_pool2d_with_indices backward does). */
int fosvos_maxpool2x2_bwd(const void* x, const void* dy, void* dx, int N, int H, int W, int C,
                          int dtype, fosvos_stream_t stream);
/* Same with gradient fan-in: dx = add + pool gradient (add: same shape as dx, may alias it, may be NULL). */
int fosvos_maxpool2x2_bwd_add(const void* x, const void* dy, const void* add, void* dx, int N, int H,
                              int W, int C, int dtype, fosvos_stream_t stream);
/* Same from the recorded window indices (fosvos_conv3x3_tc_pool_arg) instead of the pool's input: the
 * full-resolution activation is not read.  C % 32 == 0. */
int fosvos_maxpool2x2_bwd_arg(const void* pool_arg, const void* dy, const void* add, void* dx, int N,
                              int H, int W, int C, int dtype, fosvos_stream_t stream);

/* ---- side-output chain (osvos_vgg.py:69-82) -------------------------------------------
 * score_dsn 1x1 (:75) -> upscale_ ConvT (:76) -> crop (:77) for the four side maps, and
 * upscale ConvT (:71) -> crop (:72) -> cat (:80) -> fuse 1x1 (:81) for the fused map, plus the
 * consumer's sigmoid (util/experiment_helper.py:57) and 0.5 threshold (run_webcam.py:92-93).
 *
 * fosvos_side_params: device-resident parameter block, built by fosvos_side_prepare from the
 * reference-layout weights.  It holds, for each stage i=0..3 (stride s=2^(i+1), k=2s):
 *   sw[16], sb        score_dsn[i].weight/.bias
 *   g_[k*k]           upscale_[i].weight (1,1,k,k)
 *   G[k*k*16]         effective fused kernel  G[ky][kx][c] = sum_co fuse.w[16i+co] * upscale[i].w[c,co,ky,kx]
 * and fuse.bias.  This is exact for ANY upscale weights (no diagonality assumption).
 * fosvos_side_params_bytes() gives the block size. */
size_t fosvos_side_params_bytes(void);
/* Float index inside the parameter block of the "not separable" counter written by fosvos_side_prepare:
 * 0 <=> both shared up-sampling kernels factor exactly as a[ky]*b[kx] and fosvos_side_fwd may run with
 * general = 2 (two packed FMAs per stage and pixel).  general: 0 = shared-kernel fast path, 1 = any
 * up-sampling weights, 2 = separable fast path. */
int fosvos_side_params_separable_flag(void);
int fosvos_side_prepare(const float* const* upscale_w /*4: (16,16,k,k)*/,
                        const float* const* upscale1_w /*4: (1,1,k,k)*/,
                        const float* const* score_w /*4: (1,16,1,1)*/,
                        const float* const* score_b /*4: (1)*/, const float* fuse_w /*(1,64,1,1)*/,
                        const float* fuse_b /*(1)*/, void* params, fosvos_stream_t stream);
/* sp[i]: side_prep output of stage i+1, NHWC (N,h_i,w_i,16) `dtype`.  All pointer arrays
 * (`sp`, `out`, ...) are HOST arrays of device pointers.
 * out[0..3]: side maps, out[4]: fused map, each NCHW fp32 (N,1,H,W) logits (may not be NULL).
 * prob (fp32) / mask (uint8) of the fused map: optional, NULL to skip.
 * general=0: fast path, valid when fosvos_side_check_diagonal reports 0 violations; needs a
 *            workspace of fosvos_side_workspace_bytes(h,w,N) bytes (low-res head maps).
 * general=2: the same for up-sampling kernels that factor exactly into a column and a row
 *            vector (params[fosvos_side_params_separable_flag()] == 0): two FMAs per stage and pixel.
 * general=1: exact for arbitrary upscale weights (16x more arithmetic), workspace unused. */
size_t fosvos_side_workspace_bytes(const int* h /*4*/, const int* w /*4*/, int N);
/* Launch plan of the general=2 kernel for even W on a GPU with `num_sms` SMs: rows and pixel pairs per work item
 * and the dynamic shared memory of a block (two staging buffers of low-res taps).  Host arithmetic only (no
 * device needed): exposed so that the staging-window bounds can be checked on a CPU-only machine. */
int fosvos_side_upsample_plan(int N, int H, int W, int num_sms, int* rows_per_item, int* pairs_per_item,
                              int* smem_bytes);
int fosvos_side_fwd(const void* const* sp /*4*/, const int* h /*4*/, const int* w /*4*/,
                    const void* params, float* const* out /*5*/, float* prob, uint8_t* mask,
                    void* workspace, int general, int N, int H, int W, int dtype,
                    fosvos_stream_t stream);
/* The same as fosvos_side_fwd with general = 0 / 2 when the low-res head maps already sit in `workspace`
 * (stage after stage, (N,h_i,w_i) float2 each: what fosvos_conv3x3_side_tc writes through `zs`): the heads launch
 * is skipped and the side_prep maps themselves are not read.  fosvos_side_params_heads_offset(i): float index of
 * stage i's 36 head parameters {score_dsn.w[16], score_dsn.b, 3 unused, fuse.w[16 i .. 16 i + 15]} in the block. */
int fosvos_side_params_heads_offset(int stage);
int fosvos_side_fwd_heads_done(const int* h /*4*/, const int* w /*4*/, const void* params,
                               float* const* out /*5*/, float* prob, uint8_t* mask, const void* workspace,
                               int general, int N, int H, int W, fosvos_stream_t stream);
/* Backward of the chain for upscale weights that are diagonal with one shared k x k kernel per
 * stage (what interp_surgery builds, osvos_layers.py:70-81, and what lr=0 keeps,
 * network_provider.py:154-155).  dout[4] = d fused (required), dout[0..3] = d side maps or NULL.
 * Writes dsp[i] (N,h_i,w_i,16) `dtype` and accumulates (+=) into the fp32 parameter gradients
 * d_fuse_w[64], d_fuse_b[1], d_score_w[i][16], d_score_b[i][1] (NULL to skip). */
int fosvos_side_bwd(const void* const* sp, const int* h, const int* w, const void* params,
                    const float* const* dout /*5*/, void* const* dsp /*4*/, float* d_fuse_w,
                    float* d_fuse_b, float* const* d_score_w, float* const* d_score_b, int N, int H,
                    int W, int dtype, fosvos_stream_t stream);
/* Counts the elements of upscale[0..3].weight that break "diagonal with one shared k x k kernel"
 * (0 = the fast forward path and the backward are valid).  Result: int32 at *violations_dev
 * (device); no sync. */
int fosvos_side_check_diagonal(const float* const* upscale_w, int* violations_dev, fosvos_stream_t stream);

/* ---- class-balanced sigmoid cross entropy (layers/osvos_layers.py:17-44) -------------
 * stats (device, 8 doubles): [0]=#pos labels, [1]=#neg, [2]=sum_{y=1} -lv, [3]=sum_{y=0} -lv,
 * [4..7] scratch.
 * loss (device fp32 scalar) = neg/total*stats[2] + pos/total*stats[3], / numel if size_average.
 * fwd zeroes and fills stats; bwd reads it:  dx = g * w(y) * (sigmoid(x) - y) [/ numel],
 * g = *grad_out (device fp32 scalar, NULL -> 1) times grad_scale. */
/* `stats`: fosvos_bal_loss_stats_bytes() bytes of doubles: [0] positives, [1] negatives, [2] sum loss_pos,
 * [3] sum loss_neg, [4] block counter (left at 0), [8..] per-block partial sums. */
size_t fosvos_bal_loss_stats_bytes(void);
int fosvos_bal_loss_fwd(const float* output, const float* label, long long numel, int size_average,
                        double* stats, float* loss, fosvos_stream_t stream);
/* Forward and backward in ONE pass over output/label (12 B/pixel): needs the label counts up front --
 * `stats` comes from an earlier fosvos_bal_loss_fwd on the SAME label (one-shot fine-tuning keeps the label for
 * hundreds of iterations, train_online.py:77-81).  dx = grad_scale * (*grad_out or 1) * d loss / d output. */
int fosvos_bal_loss_fwd_bwd(const float* output, const float* label, long long numel, int size_average,
                            double* stats, float* loss, const float* grad_out, float grad_scale,
                            float* dx, fosvos_stream_t stream);
/* The same for n_frames maps in one launch, each frame with ITS OWN label statistics (the class balance is per
 * frame, osvos_layers.py:26-39): frame f uses output/label/dx + f*numel_per_frame, stats + f*stats_stride
 * (doubles) and loss[f]. */
int fosvos_bal_loss_fwd_bwd_frames(const float* output, const float* label, long long numel_per_frame,
                                   int n_frames, int size_average, double* stats, long long stats_stride,
                                   float* loss, const float* grad_out, float grad_scale, float* dx,
                                   fosvos_stream_t stream);
int fosvos_bal_loss_bwd(const float* output, const float* label, long long numel, int size_average,
                        const double* stats, const float* grad_out, float grad_scale, float* dx,
                        fosvos_stream_t stream);
/* Loss bookkeeping of the loops (train_online.py:82,98; train_offline.py:84-88 deep-supervision weighting), on the device:
 * total[i] = (init ? 0 : total[i]) + (*weight or 1) * part[i], i < n;   then   *loss_sum += sum(total), *last_loss = total[n-1]. */
int fosvos_loss_accumulate(const float* part, const float* weight, float* total, int n, int init,
                           fosvos_stream_t stream);
int fosvos_loss_window_finish(const float* total, int n, float* loss_sum, float* last_loss,
                              fosvos_stream_t stream);

/* ---- SGD with momentum, multi-tensor (torch.optim.SGD.step, train_online.py:99;
 *      groups: util/network_provider.py:144-159) ---------------------------------------
 * One launch updates every tensor:  d = g + wd*p;  buf = mu*buf + d;  p -= lr*buf.
 * (buf zero-initialised == torch's "buf = d on the first step".)
 * `table` is a device array of n_tensors fosvos_sgd_entry. zero_grad!=0 also clears g. */
typedef struct fosvos_sgd_entry {
  float* p;
  float* g;
  float* buf;
  long long n;
  float lr;
  float weight_decay;
} fosvos_sgd_entry;
int fosvos_sgd_chunk_elems(void); /* elements per work item: chunk_prefix[t] = sum_{u<t} ceil(n_u / this) */
int fosvos_sgd_step(const fosvos_sgd_entry* table, int n_tensors, const long long* chunk_prefix,
                    int n_chunks, float momentum, int zero_grad, fosvos_stream_t stream);

/* ---- optimizer-step companions of the tensor-core path: ONE launch each for all conv layers ----
 * fold: for every entry  dw (OIHW fp32 .grad) += ws ([tap][M][N] accumulator of
 * fosvos_conv3x3_wgrad_tc_accumulate);  ws = 0.   x_is_a = fosvos_conv3x3_wgrad_tc_orientation().
 * repack: rebuild the packed bf16 copies of the VALID (co < Cout, ci < Cin) region -- the padding of
 * buffers first filled by fosvos_pack_conv3x3_weight stays zero -- and the padded fp32 bias.
 * `tile_prefix[e]` = sum over earlier entries of fosvos_{fold,repack}_tile_count(Cout, Cin); n_entries + 1 ints. */
typedef struct fosvos_fold_entry {
  float* ws;
  float* dw;
  int Cout, Cin, CoutP, CinP;
  int x_is_a;
  int pad_;
} fosvos_fold_entry;
typedef struct fosvos_repack_entry {
  const float* w;      /* (Cout,Cin,3,3) fp32 */
  const float* bias;   /* Cout fp32 or NULL */
  void* out_fwd;       /* FOSVOS_W_TC_FWD bf16 or NULL */
  void* out_dgrad;     /* FOSVOS_W_TC_DGRAD bf16 or NULL */
  float* bias_out;     /* CoutP fp32 or NULL */
  int Cout, Cin;
  int pad_ci, pad_co;  /* K extents of the two packed layouts: ceil64(CinP), ceil64(CoutP) */
} fosvos_repack_entry;
/* fold + SGD + repack of the 3x3 conv weights fused into ONE launch (the single-GPU fine-tune step):
 * g = ws (+ dw if non-NULL; both cleared);  buf = momentum*buf + (g + weight_decay*w);  w -= lr*buf;  packed copies and
 * padded bias rebuilt from the new w (bias itself is stepped by fosvos_sgd_step before).  Tiles as fosvos_repack_tile_count. */
typedef struct fosvos_convstep_entry {
  float* ws;           /* [tap][M][N] accumulator of fosvos_conv3x3_wgrad_tc_accumulate */
  float* dw;           /* OIHW .grad or NULL */
  float* w;            /* (Cout,Cin,3,3) fp32 parameter */
  float* buf;          /* momentum buffer, same shape */
  const float* bias;   /* Cout fp32 or NULL */
  void* out_fwd;       /* FOSVOS_W_TC_FWD bf16 or NULL */
  void* out_dgrad;     /* FOSVOS_W_TC_DGRAD bf16 or NULL */
  float* bias_out;     /* CoutP fp32 or NULL */
  int Cout, Cin, CoutP, CinP;
  int x_is_a, pad_ci, pad_co, pad_;
  float lr, weight_decay;
} fosvos_convstep_entry;
int fosvos_conv_step_all(const fosvos_convstep_entry* table, int n_entries, const int* tile_prefix, int n_tiles,
                         float momentum, fosvos_stream_t stream);
int fosvos_conv3x3_wgrad_tc_orientation(int CinP, int CoutP);
int fosvos_fold_tile_count(int Cout, int Cin);
int fosvos_repack_tile_count(int Cout, int Cin);
int fosvos_wgrad_fold_all(const fosvos_fold_entry* table, int n_entries, const int* tile_prefix,
                          int n_tiles, fosvos_stream_t stream);
int fosvos_repack_all(const fosvos_repack_entry* table, int n_entries, const int* tile_prefix,
                      int n_tiles, fosvos_stream_t stream);

/* ---- callers either side of the path ----------------------------------------
 * Adam with torch.optim.Adam semantics (mimic.py:74, prune.py fine_tune): g += wd*p; m,v EMA; bias-corrected
 * update.  `state` = one device int64 step counter, incremented by the call (CUDA-graph friendly). */
typedef struct fosvos_adam_entry {
  float* p;
  float* g;
  float* m;
  float* v;
  long long n;
  float lr;
  float weight_decay;
} fosvos_adam_entry;
int fosvos_adam_chunk_elems(void);
int fosvos_adam_step(const fosvos_adam_entry* table, int n_tensors, const long long* chunk_prefix,
                     int n_chunks, float beta1, float beta2, float eps, long long* state,
                     int zero_grad, fosvos_stream_t stream);
/* nn.MSELoss (kind 0) / nn.L1Loss (kind 1) of two fp32 maps, forward and gradient in one pass
 * (mimic.py:76-81): *loss = sum or mean;  dx (may be NULL) = grad_scale * d loss / d output. */
int fosvos_pixel_loss(const float* output, const float* target, long long numel, int kind,
                      int size_average, float grad_scale, float* loss, float* dx,
                      fosvos_stream_t stream);
/* nn.ReLU(inplace=True) (osvos_vgg.py:93) at module granularity: the stage convs fuse it into their epilogue; these
 * two serve the introspection path, where a leaf module is called on its own (hooks, prune.py:94-103).
 * y = max(x, 0);  dx = dy where y > 0 else 0.  fp32, any layout. */
int fosvos_relu_fwd(const float* x, float* y, long long numel, fosvos_stream_t stream);
int fosvos_relu_bwd(const float* y, const float* dy, float* dx, long long numel, fosvos_stream_t stream);
/* Taylor pruning criterion (prune.py:163-178): rank[c] += sum_p act[p,c]*grad[p,c] / (N*H*W). NHWC. */
int fosvos_taylor_rank(const void* act, const void* grad, float* rank, int N, int H, int W, int CP,
                       int C, int dtype, fosvos_stream_t stream);
/* Frame ingest (dataloaders/davis_2016.py:115-128 + ToTensor): device uint8 (N,H,W,3) as cv2 delivers it
 * -> (N,H,W,8) activations, channel c = img[c] - mean3[c] (mean3: 3 HOST floats), channels 3..7 zero. */
int fosvos_ingest_u8(const uint8_t* img, void* y_nhwc8, int N, int H, int W, const float* mean3,
                     int dtype, fosvos_stream_t stream);

/* ---- mask egress --------------------------------------------------------------------
 * counts (device, 2 x int64 per frame): intersection and union pixel counts of two uint8
 * {0,1} masks -- the integers of the DAVIS J (region IoU) measure. Zeroed by the call. */
int fosvos_mask_iou(const uint8_t* a, const uint8_t* b, long long pixels_per_frame, int n_frames,
                    long long* counts, fosvos_stream_t stream);
/* logits fp32 -> probability fp32 and/or uint8 mask (either may be NULL). */
int fosvos_sigmoid_threshold(const float* logits, float* prob, uint8_t* mask, long long numel,
                             fosvos_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FOSVOS_B200_H_ */
