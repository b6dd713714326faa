// Multi-tensor SGD with momentum and weight decay: one launch for every parameter
// (torch.optim.SGD.step over the groups of reference util/network_provider.py:144-159,
//  called at train_online.py:99; 40 tensors -> 40+ launches there).
// HBM-bound: 12 B read + 8 B written per parameter (+4 B when zeroing the gradient).
#include "common.cuh"

namespace fosvos {

constexpr int SGD_CHUNK = 4096;   // elements per CTA work item

__global__ void __launch_bounds__(256)
sgd_kernel(const fosvos_sgd_entry* __restrict__ table, int n_tensors, const long long* __restrict__ prefix,
           int n_chunks, float mu, int zero_grad) {
  for (int chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    // binary search: tensor t with prefix[t] <= chunk < prefix[t+1]
    int lo = 0, hi = n_tensors - 1;
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= chunk) lo = mid; else hi = mid - 1;
    }
    const fosvos_sgd_entry e = table[lo];
    const long long begin = (chunk - prefix[lo]) * (long long)SGD_CHUNK;
    const long long end = min(e.n, begin + SGD_CHUNK);
    const float lr = e.lr, wd = e.weight_decay;
    float* __restrict__ p = e.p; float* __restrict__ g = e.g; float* __restrict__ b = e.buf;
    const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)b) & 15) == 0;
    const long long nvec = vec ? (end - begin) / 4 : 0;
    for (long long v = threadIdx.x; v < nvec; v += 256) {
      const long long i = begin + 4 * v;
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 bv = *reinterpret_cast<float4*>(b + i);
      bv.x = mu * bv.x + (gv.x + wd * pv.x); pv.x -= lr * bv.x;
      bv.y = mu * bv.y + (gv.y + wd * pv.y); pv.y -= lr * bv.y;
      bv.z = mu * bv.z + (gv.z + wd * pv.z); pv.z -= lr * bv.z;
      bv.w = mu * bv.w + (gv.w + wd * pv.w); pv.w -= lr * bv.w;
      *reinterpret_cast<float4*>(b + i) = bv;
      *reinterpret_cast<float4*>(p + i) = pv;
      if (zero_grad) *reinterpret_cast<float4*>(g + i) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (long long j = begin + nvec * 4 + threadIdx.x; j < end; j += 256) {   // tail / unaligned tensors
      const float gv = g[j], pv = p[j];
      const float bv = mu * b[j] + (gv + wd * pv);
      b[j] = bv;
      p[j] = pv - lr * bv;
      if (zero_grad) g[j] = 0.f;
    }
  }
}

}  // namespace fosvos

using namespace fosvos;

extern "C" {

int fosvos_sgd_chunk_elems(void) { return SGD_CHUNK; }

int fosvos_sgd_step(const fosvos_sgd_entry* table, int n_tensors, const long long* chunk_prefix, int n_chunks,
                    float momentum, int zero_grad, fosvos_stream_t stream) {
  FOSVOS_REQUIRE(table && chunk_prefix && n_tensors > 0 && n_chunks > 0, "sgd_step: bad arguments");
  const int blocks = min(n_chunks, num_sms() * 8);
  sgd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(table, n_tensors, chunk_prefix, n_chunks, momentum, zero_grad);
  return check_launch("sgd_step");
}

}  // extern "C"
