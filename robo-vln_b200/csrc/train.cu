// Training-step tail of the DAgger update (robo_vln_baselines/hierarchical_trainer.py:492-560), SURVEY.md 8(f) rank 3:
//  * hi_loss:   logits.masked_fill_(oracle == 0, 0); CrossEntropyLoss(ignore_index=-1, mean)(logits, oracle - 1)   (:506-513)
//  * lo_loss:   actions.masked_fill_(corrected == 0, 0); MSELoss(mean) + BCEWithLogitsLoss over oracle_stop != -1  (:539-553)
//    both as ONE launch that produces the loss value(s) AND the gradient w.r.t. the model outputs (the reference runs
//    ~25 elementwise / reduction kernels per loss, forward and backward);
//  * fused Adam / AdamW: ONE launch updates every trainable tensor of a model (torch.optim.AdamW for hi, torch.optim.Adam
//    with L2 weight decay for lo, :329-334), same arithmetic as torch's single-tensor path.
#include "common.cuh"
#include "rvb.h"

#include <algorithm>

namespace rvb {

namespace {

RVB_DEVICE float block_sum(float v, float* scratch) {   // all threads return the total; blockDim.x <= 1024
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float t = 0.0f;
  for (int i = 0; i < nw; ++i) t += scratch[i];   // fixed order
  return t;
}

// loss_out[0] = mean CE over rows with oracle != 0, loss_out[1] = number of such rows; dlogits = d loss / d logits
// (upstream gradient 1).  One CTA: T is at most a few thousand rows (tbptt chunk x batch).
__global__ void __launch_bounds__(256) hi_loss_kernel(const float* __restrict__ logits, const float* __restrict__ oracle_f,
                                                      const long long* __restrict__ oracle_i, int T, int C,
                                                      float* __restrict__ loss_out, float* __restrict__ dlogits) {
  __shared__ float scratch[32];
  float nll = 0.0f, cnt = 0.0f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const long long o = oracle_i != nullptr ? oracle_i[t] : static_cast<long long>(oracle_f[t]);
    if (o != 0) {             // target o - 1 >= 0; rows with o == 0 have target -1 = ignore_index (and zeroed logits)
      float mx = -INFINITY;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[t * C + c]);
      float se = 0.0f;
      for (int c = 0; c < C; ++c) se += expf(logits[t * C + c] - mx);
      nll += (mx + logf(se)) - logits[t * C + (o - 1)];
      cnt += 1.0f;
    }
  }
  const float tot = block_sum(nll, scratch);
  const float n = block_sum(cnt, scratch);
  if (threadIdx.x == 0) {
    loss_out[0] = tot / n;    // n == 0 -> NaN, as torch's mean over an empty selection
    loss_out[1] = n;
  }
  if (dlogits == nullptr) return;
  const float inv = 1.0f / n;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const long long o = oracle_i != nullptr ? oracle_i[t] : static_cast<long long>(oracle_f[t]);
    if (o == 0) {
      for (int c = 0; c < C; ++c) dlogits[t * C + c] = 0.0f;   // masked_fill_ cuts the gradient; ignored by the loss anyway
      continue;
    }
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[t * C + c]);
    float se = 0.0f;
    for (int c = 0; c < C; ++c) se += expf(logits[t * C + c] - mx);
    for (int c = 0; c < C; ++c) {
      const float p = expf(logits[t * C + c] - mx) / se;
      dlogits[t * C + c] = (p - (c == o - 1 ? 1.0f : 0.0f)) * inv;
    }
  }
}

// loss_out[0] = MSE(mean over T*A) of masked actions, loss_out[1] = BCE-with-logits (mean over oracle_stop != -1),
// loss_out[2] = number of valid stop rows; d_actions / d_stop = gradient of (loss_out[0] + loss_out[1]).
__global__ void __launch_bounds__(256) lo_loss_kernel(const float* __restrict__ actions, const float* __restrict__ corrected,
                                                      const float* __restrict__ stop, const float* __restrict__ oracle_stop,
                                                      int T, int A, float* __restrict__ loss_out, float* __restrict__ d_actions,
                                                      float* __restrict__ d_stop) {
  __shared__ float scratch[32];
  float se = 0.0f, bce = 0.0f, cnt = 0.0f;
  for (int i = threadIdx.x; i < T * A; i += blockDim.x) {
    const float y = corrected[i];
    const float x = (y == 0.0f) ? 0.0f : actions[i];   // output.masked_fill_(corrected == 0, 0)
    se += (x - y) * (x - y);
  }
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float y = oracle_stop[t];
    if (y != -1.0f) {
      const float x = stop[t];
      bce += fmaxf(x, 0.0f) - x * y + log1pf(expf(-fabsf(x)));   // numerically stable BCE with logits
      cnt += 1.0f;
    }
  }
  const float s0 = block_sum(se, scratch), s1 = block_sum(bce, scratch), n = block_sum(cnt, scratch);
  if (threadIdx.x == 0) {
    loss_out[0] = s0 / static_cast<float>(T * A);
    loss_out[1] = s1 / n;
    loss_out[2] = n;
  }
  if (d_actions != nullptr) {
    const float k = 2.0f / static_cast<float>(T * A);
    for (int i = threadIdx.x; i < T * A; i += blockDim.x) {
      const float y = corrected[i];
      d_actions[i] = (y == 0.0f) ? 0.0f : k * (actions[i] - y);
    }
  }
  if (d_stop != nullptr) {
    const float inv = 1.0f / n;
    for (int t = threadIdx.x; t < T; t += blockDim.x) {
      const float y = oracle_stop[t];
      d_stop[t] = (y != -1.0f) ? (1.0f / (1.0f + expf(-stop[t])) - y) * inv : 0.0f;
    }
  }
}

struct AdamList {
  float* const* param;
  const float* const* grad;
  float* const* exp_avg;
  float* const* exp_avg_sq;
  const long long* numel;
  const long long* chunk_start;   // [n + 1] prefix sum of ceil(numel / chunk) -> which tensor a CTA works on
};

constexpr int ADAM_CHUNK = 4096;   // elements per CTA

// torch/optim/adam.py _single_tensor_adam (amsgrad = False, maximize = False), fp32:
//   decoupled (AdamW): p *= 1 - lr*wd      else (Adam): g += wd * p
//   m = lerp(m, g, 1 - b1);  v = v*b2 + (1 - b2) g*g;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
__global__ void __launch_bounds__(256) fused_adam_kernel(AdamList L, int n_tensors, float decay_mul, float omb1, float b2, float omb2,
                                                         float eps, float wd, int decoupled, float step_size, float bc2_sqrt) {
  // binary search: tensor whose chunk range holds blockIdx.x
  int lo = 0, hi = n_tensors;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (L.chunk_start[mid] <= static_cast<long long>(blockIdx.x)) lo = mid; else hi = mid;
  }
  const int t = lo;
  const long long base = (static_cast<long long>(blockIdx.x) - L.chunk_start[t]) * ADAM_CHUNK;
  const long long n = L.numel[t];
  float* __restrict__ p = L.param[t];
  const float* __restrict__ g = L.grad[t];
  float* __restrict__ m = L.exp_avg[t];
  float* __restrict__ v = L.exp_avg_sq[t];
  for (long long i = base + threadIdx.x; i < std::min<long long>(base + ADAM_CHUNK, n); i += blockDim.x) {
    float pi = p[i], gi = g[i];
    if (decoupled) pi *= decay_mul;
    else if (wd != 0.0f) gi = fmaf(wd, pi, gi);
    const float mi = m[i] + omb1 * (gi - m[i]);
    const float vi = v[i] * b2 + omb2 * (gi * gi);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  }
}

}  // namespace

void hi_loss(const float* logits, const float* oracle_f, const int64_t* oracle_i, int T, int C, float* loss_out, float* dlogits,
             cudaStream_t s) {
  RVB_CHECK(T >= 1 && C >= 1 && C <= 64 && (oracle_f != nullptr) != (oracle_i != nullptr), "hi_loss: bad arguments");
  launch_k(hi_loss_kernel, dim3(1), dim3(256), 0, s, logits, oracle_f, reinterpret_cast<const long long*>(oracle_i), T, C, loss_out, dlogits);
  RVB_CUDA(cudaGetLastError());
}

void lo_loss(const float* actions, const float* corrected, const float* stop, const float* oracle_stop, int T, int A, float* loss_out,
             float* d_actions, float* d_stop, cudaStream_t s) {
  RVB_CHECK(T >= 1 && A >= 1, "lo_loss: bad arguments");
  launch_k(lo_loss_kernel, dim3(1), dim3(256), 0, s, actions, corrected, stop, oracle_stop, T, A, loss_out, d_actions, d_stop);
  RVB_CUDA(cudaGetLastError());
}

void fused_adam(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                const long long* numel, const long long* chunk_start, int n_tensors, long long total_chunks, double lr, double b1,
                double b2, double eps, double wd, int decoupled, double step_size, double bc2_sqrt, cudaStream_t s) {
  RVB_CHECK(n_tensors >= 1 && total_chunks >= 1 && total_chunks < (1ll << 31), "fused_adam: bad arguments");
  AdamList L{params, grads, exp_avg, exp_avg_sq, numel, chunk_start};
  // scalar constants are formed in double precision and rounded once, exactly as torch passes Python floats to its kernels
  launch_k(fused_adam_kernel, dim3(static_cast<unsigned>(total_chunks)), dim3(256), 0, s, L, n_tensors, static_cast<float>(1.0 - lr * wd),
           static_cast<float>(1.0 - b1), static_cast<float>(b2), static_cast<float>(1.0 - b2), static_cast<float>(eps),
           static_cast<float>(wd), decoupled, static_cast<float>(step_size), static_cast<float>(bc2_sqrt));
  RVB_CUDA(cudaGetLastError());
}

int adam_chunk_elems() { return ADAM_CHUNK; }

}  // namespace rvb
