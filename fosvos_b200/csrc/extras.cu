// Kernels of the callers either side of the hot path, all HBM-bound:
//   adam_kernel        multi-tensor Adam, torch.optim.Adam semantics (reference mimic.py:74, prune.py fine_tune)
//   pixel_loss_kernel  nn.MSELoss / nn.L1Loss forward + gradient in one pass (mimic.py:76-81 criteria)
//   taylor_rank_kernel per-channel sum of activation * gradient (prune.py:163-178 compute_rank)
//   ingest_u8_kernel   uint8 HWC frame - mean -> NHWC8 (dataloaders/davis_2016.py:127-128 + ToTensor)
#include "common.cuh"

namespace fosvos {

constexpr int ADAM_CHUNK = 4096;

// state[0] = step count (incremented by adam_advance_kernel before every update)
__global__ void adam_advance_kernel(long long* __restrict__ state) { state[0] += 1; }

__global__ void __launch_bounds__(256)
adam_kernel(const fosvos_adam_entry* __restrict__ table, int n_tensors, const long long* __restrict__ prefix, int n_chunks,
            float beta1, float beta2, float eps, const long long* __restrict__ state, int zero_grad) {
  const float step = (float)state[0];
  const float bc1 = 1.f - powf(beta1, step), bc2_sqrt = sqrtf(1.f - powf(beta2, step));
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_tensors - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const fosvos_adam_entry e = table[lo];
    const long long begin = (chunk - prefix[lo]) * (long long)ADAM_CHUNK;
    const long long end = min(e.n, begin + ADAM_CHUNK);
    const float step_size = e.lr / bc1;
    for (long long j = begin + threadIdx.x; j < end; j += 256) {
      const float p = e.p[j];
      const float g = e.g[j] + e.weight_decay * p;            // L2 penalty folded into the gradient (torch Adam)
      const float m = beta1 * e.m[j] + (1.f - beta1) * g;
      const float v = beta2 * e.v[j] + (1.f - beta2) * g * g;
      e.m[j] = m;
      e.v[j] = v;
      e.p[j] = p - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
      if (zero_grad) e.g[j] = 0.f;
    }
  }
}

// kind 0: squared error, kind 1: absolute error.  loss += sum (or mean); dx = scale * d loss / d x
__global__ void __launch_bounds__(256)
pixel_loss_kernel(const float* __restrict__ x, const float* __restrict__ t, long long n, int kind, float norm,
                  float scale, float* __restrict__ loss, float* __restrict__ dx) {
  float acc = 0.f;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = x[i] - t[i];
    float g;
    if (kind == 0) { acc += d * d; g = 2.f * d; }
    else { acc += fabsf(d); g = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f); }
    if (dx) dx[i] = scale * norm * g;
  }
  acc = warp_sum(acc);
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    acc = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
    acc = warp_sum(acc);
    if (threadIdx.x == 0) atomicAdd(loss, acc * norm);
  }
}

// rank[c] += inv_norm * sum_p act[p][c] * grad[p][c]      (NHWC, CP % 8 == 0)
template <typename T>
__global__ void __launch_bounds__(256)
taylor_rank_kernel(const T* __restrict__ act, const T* __restrict__ grad, float* __restrict__ rank, long long pixels, int CP,
                   int C, float inv_norm) {
  const int groups = CP / 8;
  const int g = threadIdx.x % groups;               // blockDim.x is a multiple of groups (host guarantees)
  const int lanes = blockDim.x / groups;
  const int pl = threadIdx.x / groups;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (long long px = (long long)blockIdx.x * lanes + pl; px < pixels; px += (long long)gridDim.x * lanes) {
    float a[8], b[8];
    load8(act + px * CP + g * 8, a);
    load8(grad + px * CP + g * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[j], b[j], acc[j]);
  }
  __shared__ float sm[512];
  for (int i = threadIdx.x; i < CP; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) atomicAdd(&sm[g * 8 + j], acc[j]);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(rank + i, sm[i] * inv_norm);
}

// (N,H,W,3) uint8 -> (N,H,W,8): channel c = img[c] - mean[c], channels 3..7 = 0
template <typename T>
__global__ void __launch_bounds__(256)
ingest_u8_kernel(const uint8_t* __restrict__ img, T* __restrict__ y, long long pixels, float m0, float m1, float m2) {
  for (long long p = blockIdx.x * (long long)blockDim.x + threadIdx.x; p < pixels; p += (long long)gridDim.x * blockDim.x) {
    const uint8_t* s = img + 3 * p;
    float v[8] = {(float)s[0] - m0, (float)s[1] - m1, (float)s[2] - m2, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8(y + 8 * p, v);
  }
}

}  // namespace fosvos

using namespace fosvos;

// nn.ReLU at module granularity (introspection path: a leaf module called on its own, osvos_vgg.py:93)
__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] = fmaxf(x[i], 0.f);
}
__global__ void __launch_bounds__(256) relu_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, float* __restrict__ dx, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

extern "C" {

int fosvos_relu_fwd(const float* x, float* y, long long numel, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(x && y && numel > 0, "relu_fwd: bad arguments");
  relu_fwd_kernel<<<(int)min((long long)num_sms() * 8, ceil_div_ll(numel, 256)), 256, 0, as_stream(stream)>>>(x, y, numel);
  return check_launch("relu_fwd");
}

int fosvos_relu_bwd(const float* y, const float* dy, float* dx, long long numel, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(y && dy && dx && numel > 0, "relu_bwd: bad arguments");
  relu_bwd_kernel<<<(int)min((long long)num_sms() * 8, ceil_div_ll(numel, 256)), 256, 0, as_stream(stream)>>>(y, dy, dx, numel);
  return check_launch("relu_bwd");
}

int fosvos_adam_chunk_elems(void) { return ADAM_CHUNK; }

int fosvos_adam_step(const fosvos_adam_entry* table, int n_tensors, const long long* chunk_prefix, int n_chunks, float beta1,
                     float beta2, float eps, long long* state, int zero_grad, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(table && chunk_prefix && state && n_tensors > 0 && n_chunks > 0, "adam_step: bad arguments");
  adam_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(state);
  adam_kernel<<<min(n_chunks, num_sms() * 8), 256, 0, as_stream(stream)>>>(table, n_tensors, chunk_prefix, n_chunks, beta1, beta2,
                                                                          eps, state, zero_grad);
  return check_launch("adam_step");
}

int fosvos_pixel_loss(const float* output, const float* target, long long numel, int kind, int size_average, float grad_scale,
                      float* loss, float* dx, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(output && target && loss && numel > 0 && (kind == 0 || kind == 1), "pixel_loss: bad arguments");
  cudaMemsetAsync(loss, 0, sizeof(float), as_stream(stream));
  const float norm = size_average ? 1.f / (float)numel : 1.f;
  pixel_loss_kernel<<<(int)min((long long)num_sms() * 4, ceil_div_ll(numel, 1024)), 256, 0, as_stream(stream)>>>(
      output, target, numel, kind, norm, grad_scale, loss, dx);
  return check_launch("pixel_loss");
}

int fosvos_taylor_rank(const void* act, const void* grad, float* rank, int N, int H, int W, int CP, int C, int dtype,
                       fosvos_stream_t stream) {
  FOSVOS_REQUIRE(act && grad && rank && N > 0 && H > 0 && W > 0 && CP % 8 == 0 && CP > 0 && CP <= 512 && C > 0 && C <= CP,
                 "taylor_rank: bad arguments (CP=%d C=%d)", CP, C);
  const int groups = CP / 8;
  const int threads = max(groups, (256 / groups) * groups);
  const long long pixels = (long long)N * H * W;
  const int blocks = (int)min((long long)num_sms() * 4, ceil_div_ll(pixels, threads / groups));
  const float inv = 1.f / (float)pixels;                     // prune.py:171: / (N * H * W)
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    taylor_rank_kernel<T><<<blocks, threads, 0, as_stream(stream)>>>((const T*)act, (const T*)grad, rank, pixels, CP, C, inv);
  });
  return check_launch("taylor_rank");
}

int fosvos_ingest_u8(const uint8_t* img, void* y, int N, int H, int W, const float* mean3, int dtype, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(img && y && mean3 && N > 0 && H > 0 && W > 0, "ingest_u8: bad arguments");
  const long long pixels = (long long)N * H * W;
  const int blocks = (int)min((long long)num_sms() * 8, ceil_div_ll(pixels, 256));
  FOSVOS_DISPATCH_DTYPE(dtype, T, {
    ingest_u8_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(img, (T*)y, pixels, mean3[0], mean3[1], mean3[2]);
  });
  return check_launch("ingest_u8");
}

}  // extern "C"
