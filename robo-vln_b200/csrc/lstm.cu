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
// Layout of one step: 128 CTAs, each owning 4 hidden units = 16 rows of W_hh (row = unit*4 + gate),
// staged once in shared memory (16-bit, padded pitch -> conflict-free fragment loads).  The
// recurrent product gates[N,16] = (mask*h)[N,512] . W_slice^T runs on the tensor cores with
// mma.sync m16n8k16: every warp takes 16 environments, reads their fp32 h rows straight from
// global/L2 into A fragments, and feeds them as a hi + lo pair of 16-bit values (h = hi + lo to
// ~2^-20), so the state keeps fp32 accuracy over long trajectories; fp32 accumulation.  The C
// fragment layout puts (i,f) and (g,o) of one unit in neighbouring lanes: one shuffle, then the
// cell update.
#include "common.cuh"
#include "rvb.h"

namespace rvb {

namespace {

constexpr int HID = 512;
constexpr int UNITS_PER_CTA = 4;
constexpr int LS_ROWS = UNITS_PER_CTA * 4;   // 16 gate rows = two n8 tiles
constexpr int LS_WP = HID + 8;               // h16 elements per staged row (1040 B: 16-byte aligned, bank-shifted)
constexpr int LS_THREADS = 128;

RVB_DEVICE void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#if RVB_BF16
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// x = hi + lo with hi, lo representable in the 16-bit operand type
RVB_DEVICE void split_h2(float2 x, uint32_t& hi, uint32_t& lo) {
  hi = pack_h2(x.x, x.y);
  const float2 back = unpack_h2(hi);
  lo = pack_h2(x.x - back.x, x.y - back.y);
}

constexpr int LS_HP = HID + 8;               // floats per staged h row (2080 B): conflict-free float2 fragment loads
constexpr int LS_PASS = 16 * (LS_THREADS / 32);   // environments per pass (one m16 tile per warp)
constexpr int LS_SMEM = LS_ROWS * LS_WP * 2 + LS_PASS * LS_HP * 4 + 64;

__global__ void __launch_bounds__(LS_THREADS) lstm_step_kernel(const float* __restrict__ gx, const h16* __restrict__ whh,
                                                               const float* __restrict__ masks, int mask_stride,
                                                               const float* __restrict__ h_prev,
                                                               const float* c_prev, float* __restrict__ h_next,
                                                               float* c_next, float* __restrict__ h_final,
                                                               float* __restrict__ y, int t, int N) {
  extern __shared__ __align__(16) uint8_t ls_raw[];
  h16* sW = reinterpret_cast<h16*>(ls_raw);                                        // [16][LS_WP]
  float* sH = reinterpret_cast<float*>(ls_raw + LS_ROWS * LS_WP * 2);              // [LS_PASS][LS_HP]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ls_raw + LS_ROWS * LS_WP * 2 + LS_PASS * LS_HP * 4);   // one per warp
  int* s_flag = reinterpret_cast<int*>(bars + LS_THREADS / 32);
  const int u0 = blockIdx.x * UNITS_PER_CTA;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // W_hh rows of this CTA (constant): staged row r = unit_local*4 + gate
  for (int i = threadIdx.x; i < LS_ROWS * (HID / 8); i += LS_THREADS) {
    const int r = i / (HID / 8), v = i % (HID / 8);
    const int wrow = (r & 3) * HID + u0 + (r >> 2);
    *reinterpret_cast<uint4*>(sW + r * LS_WP + v * 8) =
        __ldg(reinterpret_cast<const uint4*>(whh + static_cast<long long>(wrow) * HID) + v);
  }
  if (threadIdx.x == 0) {
    *s_flag = (t == 0) ? 1 : 0;
    for (int w = 0; w < LS_THREADS / 32; ++w) mbar_init(&bars[w], 1);
    fence_barrier_init();
  }
  RVB_PDL_PROLOGUE();   // everything above touches only constant weights
  __syncthreads();
  if (t != 0 && warp == 0) {
    int any = 0;
    for (int n = lane; n < N; n += 32) any |= (masks[(static_cast<long long>(t) * N + n) * mask_stride] == 0.0f) ? 1 : 0;
    any = __any_sync(0xffffffffu, any);
    if (lane == 0 && any) *s_flag = 1;
  }
  __syncthreads();
  const bool apply_mask = *s_flag != 0;
  const int grp = lane >> 2, q = lane & 3;
  const int gate_a = (q & 1) * 2;   // even q: gates (i, f); odd q: gates (g, o)
  float* myH = sH + warp * 16 * LS_HP;
  uint32_t phase = 0;

  for (int base = 0; base < N; base += LS_PASS) {
    const int r0 = base + warp * 16;
    const int nrows = min(16, N - r0);
    if (nrows <= 0) continue;   // warp-uniform
    // this warp's 16 h rows: one 2 KB bulk copy per row, all in flight at once
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive_expect_tx(&bars[warp], static_cast<uint32_t>(nrows) * HID * 4);
    __syncwarp();
    if (lane < nrows)
      bulk_load_1d(myH + lane * LS_HP, h_prev + static_cast<long long>(r0 + lane) * HID, HID * 4, &bars[warp]);
    // while they land: masks, gx and c for this thread's outputs
    const int row[2] = {r0 + grp, r0 + grp + 8};
    float msk[2], gxa[2][2], gxb[2][2], cp[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool ok = row[i] < N;
      msk[i] = !ok ? 0.0f : (apply_mask ? masks[(static_cast<long long>(t) * N + row[i]) * mask_stride] : 1.0f);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int unit = u0 + j * 2 + (q >> 1);
        gxa[j][i] = gxb[j][i] = cp[j][i] = 0.0f;
        if (ok) {
          const float* g = gx + (static_cast<long long>(t) * N + row[i]) * (4 * HID) + unit;
          gxa[j][i] = g[gate_a * HID];
          gxb[j][i] = g[(gate_a + 1) * HID];
          if ((q & 1) == 0) cp[j][i] = c_prev[static_cast<long long>(row[i]) * HID + unit];
        }
      }
    }
    mbar_wait(&bars[warp], phase);
    phase ^= 1;

    float acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[j][e] = 0.0f;
    const float* hr[2] = {myH + grp * LS_HP, myH + (grp + 8) * LS_HP};
#pragma unroll 4
    for (int ks = 0; ks < HID / 16; ++ks) {
      const int k0 = ks * 16 + q * 2;
      uint32_t ahi[4], alo[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
#pragma unroll
        for (int hk = 0; hk < 2; ++hk) {
          float2 x = make_float2(0.0f, 0.0f);
          if (row[i] < N) {   // rows past N were never copied: do not touch the (uninitialised) slot
            x = *reinterpret_cast<const float2*>(hr[i] + k0 + hk * 8);
            x.x *= msk[i]; x.y *= msk[i];
          }
          split_h2(x, ahi[hk * 2 + i], alo[hk * 2 + i]);
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const h16* wr = sW + (j * 8 + grp) * LS_WP + k0;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(wr);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(wr + 8);
        mma_16816(acc[j], ahi, b0, b1);
        mma_16816(acc[j], alo, b0, b1);
      }
    }
    // C fragment: acc[j][0..1] = (row0, n = j*8 + 2q, +1), acc[j][2..3] = (row1, same n); n = unit_local*4 + gate
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int unit = u0 + j * 2 + (q >> 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        acc[j][2 * i] += gxa[j][i];
        acc[j][2 * i + 1] += gxb[j][i];
      }
      float other[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) other[e] = __shfl_xor_sync(0xffffffffu, acc[j][e], 1);
      if ((q & 1) == 0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (row[i] < N) {
            const float gi = acc[j][2 * i], gf = acc[j][2 * i + 1], gg = other[2 * i], go = other[2 * i + 1];
            const long long idx = static_cast<long long>(row[i]) * HID + unit;
            const float c0 = cp[j][i] * msk[i];
            const float i_ = 1.0f / (1.0f + expf(-gi));
            const float f_ = 1.0f / (1.0f + expf(-gf));
            const float o_ = 1.0f / (1.0f + expf(-go));
            const float c1 = f_ * c0 + i_ * tanhf(gg);
            const float h1 = o_ * tanhf(c1);
            c_next[idx] = c1;
            h_next[idx] = h1;
            if (h_final != nullptr) h_final[idx] = h1;
            y[(static_cast<long long>(t) * N + row[i]) * HID + unit] = h1;
          }
        }
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
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    RVB_CUDA(cudaFuncSetAttribute(lstm_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LS_SMEM));
  }
  const long long NH = static_cast<long long>(N) * HID;
  for (int t = 0; t < T; ++t) {
    const float* h_prev = (t == 0) ? hc_in : h_scratch + ((t - 1) & 1) * NH;
    const float* c_prev = (t == 0) ? hc_in + NH : hc_out + NH;
    float* h_next = h_scratch + (t & 1) * NH;
    float* h_final = (t == T - 1) ? hc_out : nullptr;
    launch_k(lstm_step_kernel, dim3(HID / UNITS_PER_CTA), dim3(LS_THREADS), LS_SMEM, s, gx, whh, masks, mask_stride, h_prev, c_prev, h_next,
                                                                 hc_out + NH, h_final, y, t, N);
  }
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
