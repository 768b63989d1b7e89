// RNNStateEncoder (LSTM, 1 layer, hidden 512) recurrence.
// Reference: habitat_baselines/rl/models/rnn_state_encoder.py:74-142 around torch.nn.LSTM
// (gate order i,f,g,o).  The input projection gx = x W_ih^T + b_ih + b_hh for all T*N rows is
// one tensor-core GEMM (gemm_tc.cu); this file does the serial part, one launch per time step:
// gates = gx[t] + (mask-reset h) W_hh^T, then the cell update.
//
// Mask semantics reproduced exactly: (h, c) are multiplied by masks[t] at t = 0 and at every
// later step where ANY env has a zero mask (the reference's segment starts); single-step
// batches (T == 1) always multiply (single_forward).
//
// Layout of one step: 64 CTAs, each owning 8 hidden units = 32 rows of W_hh (i,f,g,o x 8),
// kept in shared memory (16-bit, padded rows -> conflict-free).  Environments are processed in
// blocks of 16 whose (masked) h vectors are staged in shared memory as fp32; each warp takes
// two environments, lane r accumulates gate row r with four independent FMA chains.
#include "common.cuh"
#include "rvb.h"

namespace rvb {

namespace {

constexpr int HID = 512;
constexpr int UNITS_PER_CTA = 8;   // 8 hidden units x 4 gates = 32 rows = 32 lanes
constexpr int WPITCH = HID + 2;    // h16 elements; +2 -> row-to-row bank shift of one word
constexpr int ENV_BLOCK = 16;
constexpr int LSTM_SMEM = 32 * WPITCH * 2 + ENV_BLOCK * HID * 4 + ENV_BLOCK * 4;

__global__ void __launch_bounds__(256) lstm_step_kernel(const float* __restrict__ gx, const h16* __restrict__ whh,
                                                        const float* __restrict__ masks, int mask_stride,
                                                        const float* __restrict__ h_prev,
                                                        const float* __restrict__ c_prev, float* __restrict__ h_next,
                                                        float* __restrict__ c_next, float* __restrict__ h_final,
                                                        float* __restrict__ y, int t, int N) {
  extern __shared__ __align__(16) uint8_t sm_raw[];
  h16* sW = reinterpret_cast<h16*>(sm_raw);                                  // [32][WPITCH]
  float* sH = reinterpret_cast<float*>(sm_raw + 32 * WPITCH * 2);            // [ENV_BLOCK][HID], mask applied
  float* sM = sH + ENV_BLOCK * HID;                                          // [ENV_BLOCK] mask multipliers
  __shared__ int s_flag;
  const int u0 = blockIdx.x * UNITS_PER_CTA;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // rows of W_hh owned by this CTA: lane r <-> gate r/8, unit u0 + r%8
  for (int i = threadIdx.x; i < 32 * (HID / 8); i += blockDim.x) {
    const int r = i / (HID / 8), v = i % (HID / 8);
    const int wrow = (r >> 3) * HID + u0 + (r & 7);
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(whh + static_cast<long long>(wrow) * HID) + v);
    uint32_t* dst = reinterpret_cast<uint32_t*>(sW + r * WPITCH) + v * 4;
    dst[0] = q.x; dst[1] = q.y; dst[2] = q.z; dst[3] = q.w;
  }
  if (threadIdx.x == 0) s_flag = (t == 0) ? 1 : 0;
  RVB_PDL_PROLOGUE();   // W_hh (constant) is staged above while the previous kernel drains
  __syncthreads();
  if (t != 0 && warp == 0) {
    int any = 0;
    for (int n = lane; n < N; n += 32) any |= (masks[(static_cast<long long>(t) * N + n) * mask_stride] == 0.0f) ? 1 : 0;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0 && any) s_flag = 1;
  }
  __syncthreads();
  const bool apply_mask = s_flag != 0;
  const uint32_t* wrow = reinterpret_cast<const uint32_t*>(sW + lane * WPITCH);

  for (int nb = 0; nb < N; nb += ENV_BLOCK) {
    __syncthreads();   // previous block's sH fully consumed
    if (threadIdx.x < ENV_BLOCK) {
      const int n = nb + threadIdx.x;
      sM[threadIdx.x] = (apply_mask && n < N) ? masks[(static_cast<long long>(t) * N + n) * mask_stride] : 1.0f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ENV_BLOCK * (HID / 4); i += blockDim.x) {
      const int e = i / (HID / 4), v = i % (HID / 4);
      const int n = nb + e;
      float4 h4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) {
        h4 = *reinterpret_cast<const float4*>(h_prev + static_cast<long long>(n) * HID + v * 4);
        const float m = sM[e];
        h4.x *= m; h4.y *= m; h4.z *= m; h4.w *= m;
      }
      *reinterpret_cast<float4*>(sH + e * HID + v * 4) = h4;
    }
    __syncthreads();
    const int e0 = warp * 2, e1 = e0 + 1;
    const float2* ha = reinterpret_cast<const float2*>(sH + e0 * HID);
    const float2* hb = reinterpret_cast<const float2*>(sH + e1 * HID);
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll 8
    for (int k2 = 0; k2 < HID / 2; ++k2) {
      const float2 w2 = unpack_h2(wrow[k2]);
      const float2 x = ha[k2];
      const float2 z = hb[k2];
      a0 = fmaf(w2.x, x.x, a0);
      a1 = fmaf(w2.y, x.y, a1);
      b0 = fmaf(w2.x, z.x, b0);
      b1 = fmaf(w2.y, z.y, b1);
    }
    const int gate = lane >> 3, u = lane & 7;
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const int n = nb + (which ? e1 : e0);
      float acc = which ? (b0 + b1) : (a0 + a1);
      if (n < N) acc += gx[(static_cast<long long>(t) * N + n) * (4 * HID) + gate * HID + u0 + u];
      const float gi = __shfl_sync(0xffffffffu, acc, u);
      const float gf = __shfl_sync(0xffffffffu, acc, 8 + u);
      const float gg = __shfl_sync(0xffffffffu, acc, 16 + u);
      const float go = __shfl_sync(0xffffffffu, acc, 24 + u);
      if (lane < 8 && n < N) {
        const long long idx = static_cast<long long>(n) * HID + u0 + u;
        const float c0 = c_prev[idx] * sM[which ? e1 : e0];
        const float i_ = 1.0f / (1.0f + expf(-gi));
        const float f_ = 1.0f / (1.0f + expf(-gf));
        const float o_ = 1.0f / (1.0f + expf(-go));
        const float c1 = f_ * c0 + i_ * tanhf(gg);
        const float h1 = o_ * tanhf(c1);
        c_next[idx] = c1;
        h_next[idx] = h1;
        if (h_final != nullptr) h_final[idx] = h1;
        y[(static_cast<long long>(t) * N + n) * HID + u0 + u] = h1;
      }
    }
  }
}

}  // namespace

// gx [T*N, 2048] fp32 (biases included), whh h16 [2048, 512], masks fp32 (element (t*N+n)*mask_stride),
// hc_in / hc_out fp32 [2, N, 512] (must not alias), h_scratch fp32 [2, N, 512], y fp32 [T*N, 512].
void lstm_forward(const float* gx, const h16* whh, const float* masks, int mask_stride, const float* hc_in,
                  float* hc_out, float* h_scratch, float* y, int T, int N, cudaStream_t s) {
  RVB_CHECK(T >= 1 && N >= 1, "lstm: empty batch");
  RVB_CHECK(hc_in != hc_out, "lstm: hidden state in/out must not alias");
  static bool attr = false;
  if (!attr) {
    RVB_CUDA(cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSTM_SMEM));
    attr = true;
  }
  const long long NH = static_cast<long long>(N) * HID;
  for (int t = 0; t < T; ++t) {
    const float* h_prev = (t == 0) ? hc_in : h_scratch + ((t - 1) & 1) * NH;
    const float* c_prev = (t == 0) ? hc_in + NH : hc_out + NH;
    float* h_next = h_scratch + (t & 1) * NH;
    float* h_final = (t == T - 1) ? hc_out : nullptr;
    launch_k(lstm_step_kernel, dim3(HID / UNITS_PER_CTA), dim3(256), LSTM_SMEM, s, gx, whh, masks, mask_stride, h_prev, c_prev, h_next,
                                                                 hc_out + NH, h_final, y, t, N);
  }
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
