// Class-balanced sigmoid cross entropy, forward reduction and backward
// (reference layers/osvos_layers.py:17-44; ~15 ATen kernels + 3 full reductions there).
// HBM-bound: forward reads 8 B/pixel, backward reads 8 + writes 4 B/pixel; float4 accesses,
// warp-shuffle + one double atomic per warp-group for the sums.
#include "common.cuh"

namespace fosvos {

struct LossAcc {
  float pos, s_pos, s_neg;
};

__device__ __forceinline__ void loss_elem(float x, float lab, LossAcc& a) {
  const float y = lab >= 0.5f ? 1.f : 0.f;                 // torch.ge(label, 0.5).float()   :26
  const float z = x >= 0.f ? 1.f : 0.f;                    // output_gt_zero                :32
  const float lv = x * (y - z) - logf(1.f + expf(x - 2.f * x * z));   // loss_val           :33-34
  a.pos += y;
  a.s_pos += -(y * lv);                                    // loss_pos                      :36
  a.s_neg += -((1.f - y) * lv);                            // loss_neg                      :37
}

constexpr int LOSS_THREADS = 1024;              // the exp/log chains need the occupancy ...
constexpr int LOSS_MAX_BLOCKS = 160;            // ... but every block ends with a same-address counter atomic (~27 cycles each,
                                                // serialised): ONE fat block per SM; per-block partial sums live behind stats[8]

// Block reduction of the three sums into this block's slot; the last block to finish (one counter atomic per
// block, not three same-address double atomics) adds the slots up in a fixed order and finalises.
__device__ __forceinline__ void loss_finish(LossAcc a, double* __restrict__ stats, float* __restrict__ loss, long long n,
                                            int size_average, bool counts_given) {
  // fp32 inside the block (<= a few 10^4 elements per block: counts stay exact, sums keep ~1e-6), double across blocks
  float posf = warp_sum(a.pos), spf = warp_sum(a.s_pos), snf = warp_sum(a.s_neg);
  __shared__ float sm[3][LOSS_THREADS / 32];
  __shared__ bool last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sm[0][wid] = posf; sm[1][wid] = spf; sm[2][wid] = snf; }
  __syncthreads();
  if (wid == 0) {
    posf = lane < LOSS_THREADS / 32 ? sm[0][lane] : 0.f;
    spf = lane < LOSS_THREADS / 32 ? sm[1][lane] : 0.f;
    snf = lane < LOSS_THREADS / 32 ? sm[2][lane] : 0.f;
    posf = warp_sum(posf); spf = warp_sum(spf); snf = warp_sum(snf);
    const double pos = (double)posf, sp = (double)spf, sn = (double)snf;
    if (lane == 0) {
      double* slot = stats + 8 + 3 * blockIdx.x;
      slot[0] = pos; slot[1] = sp; slot[2] = sn;
      __threadfence();
      unsigned long long* counter = reinterpret_cast<unsigned long long*>(&stats[4]);
      last = atomicAdd(counter, 1ULL) + 1ULL == gridDim.x;
    }
  }
  __syncthreads();
  if (last) {
    // every slot in ONE round trip (a serial loop of L2 reads would cost microseconds), then a fixed-order sum
    __shared__ double part[3 * LOSS_MAX_BLOCKS];
    __threadfence();
    const int nslots = 3 * (int)gridDim.x;
    for (int t = threadIdx.x; t < nslots; t += LOSS_THREADS) part[t] = __ldcg(stats + 8 + t);
    __syncthreads();
    if (threadIdx.x < 32) {
      double P = 0.0, SP = 0.0, SN = 0.0;
      for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) { P += part[3 * b]; SP += part[3 * b + 1]; SN += part[3 * b + 2]; }
      P = warp_sum(P); SP = warp_sum(SP); SN = warp_sum(SN);
      if (threadIdx.x == 0) {
        if (counts_given) P = stats[0];
        const double Nn = (double)n - P, tot = (double)n;
        stats[0] = P;
        stats[1] = Nn;
        stats[2] = SP;
        stats[3] = SN;
        // fp32 like the reference: num_labels_neg / num_total * loss_pos + num_labels_pos / num_total * loss_neg  :39
        float v = (float)Nn / (float)tot * (float)SP + (float)P / (float)tot * (float)SN;
        if (size_average) v = v / (float)n;                                                          // :41-42
        *loss = v;
        *reinterpret_cast<unsigned long long*>(&stats[4]) = 0ULL;   // ready for the next launch (no memset node needed)
      }
    }
  }
}

__global__ void __launch_bounds__(LOSS_THREADS)
bal_loss_fwd_kernel(const float* __restrict__ out, const float* __restrict__ lab, long long n, int size_average,
                    double* __restrict__ stats, float* __restrict__ loss) {
  LossAcc a{0.f, 0.f, 0.f};
  // float4 path only for 16-byte aligned maps (a frame slice of an odd-sized batch is not)
  const long long n4 = (((uintptr_t)out | (uintptr_t)lab) & 15) == 0 ? n / 4 : 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  long long i = tid;
  for (; i + nth < n4; i += 2 * nth) {           // two independent 16-byte loads per stream in flight
    const float4 x0 = reinterpret_cast<const float4*>(out)[i], x1 = reinterpret_cast<const float4*>(out)[i + nth];
    const float4 l0 = reinterpret_cast<const float4*>(lab)[i], l1 = reinterpret_cast<const float4*>(lab)[i + nth];
    loss_elem(x0.x, l0.x, a); loss_elem(x0.y, l0.y, a); loss_elem(x0.z, l0.z, a); loss_elem(x0.w, l0.w, a);
    loss_elem(x1.x, l1.x, a); loss_elem(x1.y, l1.y, a); loss_elem(x1.z, l1.z, a); loss_elem(x1.w, l1.w, a);
  }
  for (; i < n4; i += nth) {
    const float4 x = reinterpret_cast<const float4*>(out)[i];
    const float4 l = reinterpret_cast<const float4*>(lab)[i];
    loss_elem(x.x, l.x, a); loss_elem(x.y, l.y, a); loss_elem(x.z, l.z, a); loss_elem(x.w, l.w, a);
  }
  for (long long j = n4 * 4 + tid; j < n; j += nth) loss_elem(out[j], lab[j], a);
  loss_finish(a, stats, loss, n, size_average, false);
}

// Forward AND backward in one pass (12 B/pixel instead of 20) when the label counts are known up front --
// they depend on the label only, which one-shot fine-tuning keeps for hundreds of iterations:
// stats[0] = number of positive labels on entry (from a previous forward on the same label).
__global__ void __launch_bounds__(LOSS_THREADS)
bal_loss_fused_kernel(const float* __restrict__ out, const float* __restrict__ lab, long long n, int size_average,
                      double* __restrict__ stats, float* __restrict__ loss, const float* __restrict__ grad_out, float grad_scale,
                      float* __restrict__ dx, long long stats_stride) {
  // blockIdx.y = frame: every frame has its own label statistics, partial-sum slots and loss scalar
  out += blockIdx.y * n; lab += blockIdx.y * n; dx += blockIdx.y * n;
  stats += blockIdx.y * stats_stride; loss += blockIdx.y;
  const float tot = (float)n;
  float g = grad_scale * (grad_out ? *grad_out : 1.f);
  if (size_average) g /= tot;
  const double P = stats[0];
  const float w1 = (float)((double)n - P) / tot * g;   // y = 1: neg/total
  const float w0 = (float)P / tot * g;                 // y = 0: pos/total
  LossAcc a{0.f, 0.f, 0.f};
  // One exponential serves both: t = exp(-|x|) = exp(x - 2 x [x >= 0]) (osvos_layers.py:33), loss term
  // x (y - z) - log(1 + t), sigmoid = 1 / (1 + t) or t / (1 + t).  MUFU-based exp / log / reciprocal (relative error
  // ~1e-6, far inside the 2e-5 / 1e-4 parity bounds of loss and gradient) keep this pass HBM-bound.
  auto f = [&](float x, float l) -> float {
    const float y = l >= 0.5f ? 1.f : 0.f;
    const float z = x >= 0.f ? 1.f : 0.f;
    const float t = __expf(-fabsf(x));
    const float u = 1.f + t;
    const float lv = x * (y - z) - __logf(u);
    a.pos += y;
    a.s_pos += -(y * lv);
    a.s_neg += -((1.f - y) * lv);
    const float r = __frcp_rn(u);
    const float sg = x >= 0.f ? r : t * r;
    return l >= 0.5f ? w1 * (sg - 1.f) : w0 * sg;
  };
  const long long n4 = (((uintptr_t)out | (uintptr_t)lab | (uintptr_t)dx) & 15) == 0 ? n / 4 : 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nth) {
    const float4 x = reinterpret_cast<const float4*>(out)[i];
    const float4 l = reinterpret_cast<const float4*>(lab)[i];
    reinterpret_cast<float4*>(dx)[i] = make_float4(f(x.x, l.x), f(x.y, l.y), f(x.z, l.z), f(x.w, l.w));
  }
  for (long long j = n4 * 4 + tid; j < n; j += nth) dx[j] = f(out[j], lab[j]);
  loss_finish(a, stats, loss, n, size_average, true);
}

__global__ void __launch_bounds__(256)
bal_loss_bwd_kernel(const float* __restrict__ out, const float* __restrict__ lab, long long n, int size_average,
                    const double* __restrict__ stats, const float* __restrict__ grad_out, float grad_scale,
                    float* __restrict__ dx) {
  const float tot = (float)n;
  float g = grad_scale * (grad_out ? *grad_out : 1.f);
  if (size_average) g /= tot;
  const float w1 = (float)stats[1] / tot * g;   // y = 1: neg/total
  const float w0 = (float)stats[0] / tot * g;   // y = 0: pos/total
  auto f = [&](float x, float l) -> float {
    const float t = expf(-fabsf(x));
    const float s = x >= 0.f ? 1.f / (1.f + t) : t / (1.f + t);
    return l >= 0.5f ? w1 * (s - 1.f) : w0 * s;
  };
  const long long n4 = (((uintptr_t)out | (uintptr_t)lab | (uintptr_t)dx) & 15) == 0 ? n / 4 : 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n4; i += nth) {
    const float4 x = reinterpret_cast<const float4*>(out)[i];
    const float4 l = reinterpret_cast<const float4*>(lab)[i];
    reinterpret_cast<float4*>(dx)[i] = make_float4(f(x.x, l.x), f(x.y, l.y), f(x.z, l.z), f(x.w, l.w));
  }
  for (long long i = n4 * 4 + tid; i < n; i += nth) dx[i] = f(out[i], lab[i]);
}

__global__ void sigmoid_threshold_kernel(const float* __restrict__ x, float* __restrict__ prob,
                                         uint8_t* __restrict__ mask, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float p = 1.f / (1.f + expf(-x[i]));        // util/experiment_helper.py:57
    if (prob) prob[i] = p;
    if (mask) mask[i] = p >= 0.5f ? 1 : 0;            // run_webcam.py:92-93
  }
}

// |a & b| and |a | b| of two {0,1} byte masks, one (inter, union) int64 pair per frame.
__global__ void __launch_bounds__(256)
mask_iou_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, long long ppf, long long* __restrict__ counts) {
  const int f = blockIdx.y;
  const uint8_t* pa = a + (long long)f * ppf;
  const uint8_t* pb = b + (long long)f * ppf;
  unsigned inter = 0, uni = 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(pa) | reinterpret_cast<uintptr_t>(pb)) & 15) == 0;
  const long long n16 = aligned ? ppf / 16 : 0;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  for (long long i = tid; i < n16; i += nth) {
    const uint4 va = reinterpret_cast<const uint4*>(pa)[i], vb = reinterpret_cast<const uint4*>(pb)[i];
    const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned ma = __vcmpne4(wa[k], 0u) & 0x01010101u, mb = __vcmpne4(wb[k], 0u) & 0x01010101u;
      inter += __popc(ma & mb);
      uni += __popc(ma | mb);
    }
  }
  for (long long i = n16 * 16 + tid; i < ppf; i += nth) {
    const bool x = pa[i] != 0, y = pb[i] != 0;
    inter += (x && y);
    uni += (x || y);
  }
  inter = __reduce_add_sync(0xffffffffu, inter);
  uni = __reduce_add_sync(0xffffffffu, uni);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&counts[2 * f]), (unsigned long long)inter);
    atomicAdd(reinterpret_cast<unsigned long long*>(&counts[2 * f + 1]), (unsigned long long)uni);
  }
}

// Loss bookkeeping of the fine-tune loop (train_online.py:82,98 / train_offline.py:84-88), kept on the device so that a
// captured window needs no library kernel: total[i] = (init ? 0 : total[i]) + w * part[i];  finish: sum += total[0..n), last = total[n-1]
__global__ void loss_accumulate_kernel(const float* __restrict__ part, const float* __restrict__ weight, float* __restrict__ total, int n, int init) {
  const float w = weight ? *weight : 1.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) total[i] = (init ? 0.f : total[i]) + w * part[i];
}
__global__ void loss_window_finish_kernel(const float* __restrict__ total, int n, float* __restrict__ loss_sum, float* __restrict__ last_loss) {
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += total[i];
    *loss_sum += s;
    *last_loss = total[n - 1];
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_loss_accumulate(const float* part, const float* weight, float* total, int n, int init, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(part && total && n > 0, "loss_accumulate: bad arguments");
  loss_accumulate_kernel<<<1, 64, 0, as_stream(stream)>>>(part, weight, total, n, init);
  return check_launch("loss_accumulate");
}

int fosvos_loss_window_finish(const float* total, int n, float* loss_sum, float* last_loss, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(total && loss_sum && last_loss && n > 0, "loss_window_finish: bad arguments");
  loss_window_finish_kernel<<<1, 32, 0, as_stream(stream)>>>(total, n, loss_sum, last_loss);
  return check_launch("loss_window_finish");
}

size_t fosvos_bal_loss_stats_bytes(void) { return sizeof(double) * (8 + 3 * LOSS_MAX_BLOCKS); }

static int loss_blocks(long long numel) { return (int)min((long long)min(num_sms(), LOSS_MAX_BLOCKS), ceil_div_ll(numel, 4 * LOSS_THREADS)); }

int fosvos_bal_loss_fwd(const float* output, const float* label, long long numel, int size_average, double* stats,
                        float* loss, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(output && label && stats && loss && numel > 0, "bal_loss_fwd: bad arguments");
  cudaMemsetAsync(stats, 0, 8 * sizeof(double), as_stream(stream));
  bal_loss_fwd_kernel<<<loss_blocks(numel), LOSS_THREADS, 0, as_stream(stream)>>>(output, label, numel, size_average, stats, loss);
  return check_launch("bal_loss_fwd");
}

int fosvos_bal_loss_fwd_bwd(const float* output, const float* label, long long numel, int size_average, double* stats,
                            float* loss, const float* grad_out, float grad_scale, float* dx, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(output && label && stats && loss && dx && numel > 0, "bal_loss_fwd_bwd: bad arguments");
  bal_loss_fused_kernel<<<loss_blocks(numel), LOSS_THREADS, 0, as_stream(stream)>>>(output, label, numel, size_average, stats, loss,
                                                                                   grad_out, grad_scale, dx, 0);
  return check_launch("bal_loss_fwd_bwd");
}

int fosvos_bal_loss_fwd_bwd_frames(const float* output, const float* label, long long numel_per_frame, int n_frames, int size_average,
                                   double* stats, long long stats_stride, float* loss, const float* grad_out, float grad_scale,
                                   float* dx, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(output && label && stats && loss && dx && numel_per_frame > 0 && n_frames > 0 && n_frames <= 65535,
                 "bal_loss_fwd_bwd_frames: bad arguments");
  FOSVOS_REQUIRE(stats_stride * (long long)sizeof(double) >= (long long)fosvos_bal_loss_stats_bytes(),
                 "bal_loss_fwd_bwd_frames: stats stride smaller than fosvos_bal_loss_stats_bytes()");
  const int per_frame = max(1, min(loss_blocks(numel_per_frame), ceil_div(num_sms(), n_frames)));
  dim3 grid(per_frame, n_frames);
  bal_loss_fused_kernel<<<grid, LOSS_THREADS, 0, as_stream(stream)>>>(output, label, numel_per_frame, size_average, stats, loss, grad_out,
                                                                     grad_scale, dx, stats_stride);
  return check_launch("bal_loss_fwd_bwd_frames");
}

int fosvos_bal_loss_bwd(const float* output, const float* label, long long numel, int size_average,
                        const double* stats, const float* grad_out, float grad_scale, float* dx,
                        fosvos_stream_t stream) {
  FOSVOS_REQUIRE(output && label && stats && dx && numel > 0, "bal_loss_bwd: bad arguments");
  const int blocks = (int)min((long long)num_sms() * 4, ceil_div_ll(numel, 1024));
  bal_loss_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(output, label, numel, size_average, stats, grad_out,
                                                            grad_scale, dx);
  return check_launch("bal_loss_bwd");
}

int fosvos_sigmoid_threshold(const float* logits, float* prob, uint8_t* mask, long long numel, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(logits && (prob || mask) && numel > 0, "sigmoid_threshold: bad arguments");
  const int blocks = (int)min((long long)num_sms() * 8, ceil_div_ll(numel, 256));
  sigmoid_threshold_kernel<<<blocks, 256, 0, as_stream(stream)>>>(logits, prob, mask, numel);
  return check_launch("sigmoid_threshold");
}

int fosvos_mask_iou(const uint8_t* a, const uint8_t* b, long long pixels_per_frame, int n_frames, long long* counts,
                    fosvos_stream_t stream) {
  FOSVOS_REQUIRE(a && b && counts && pixels_per_frame > 0 && n_frames > 0, "mask_iou: bad arguments");
  cudaMemsetAsync(counts, 0, 2 * sizeof(long long) * n_frames, as_stream(stream));
  dim3 grid((unsigned)min((long long)64, ceil_div_ll(pixels_per_frame, 256 * 16)), n_frames);
  mask_iou_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, pixels_per_frame, counts);
  return check_launch("mask_iou");
}

}  // extern "C"
