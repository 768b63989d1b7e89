// Attention kernels.
//  * bert_self_attention: softmax(Q K^T / 8) V per (row, head) for BERT-base (12 heads x 64),
//    all-ones mask because the reference calls BertModel with input_ids only
//    (seq2seq_highlevel_cma.py:192-195).  L <= 256 keys fit one CTA; scores stay in registers.
//    L=80 makes each head a 80x80x64 problem (1.7% of BERT's FLOPs), which does not fill a
//    128-row tcgen05 tile, so this kernel uses warp-level mma.sync m16n8k16 (one 16-query
//    tile per warp) -- the big BERT contractions go through gemm_tc.cu.
//  * vla_cross_attention: ScaledDotProductAttention of Visual_Ling_Attn
//    (transformer.py:81-109) -- 4 heads x 64 over only 16 visual keys; CUDA cores.
#include "common.cuh"
#include "rvb.h"

namespace rvb {

namespace {

constexpr int HD = 64;        // head dim
constexpr int PITCH = HD + 8; // smem row pitch (h16) -> conflict-free ldmatrix

RVB_DEVICE void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
RVB_DEVICE void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
RVB_DEVICE void mma_h16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                               uint32_t b1) {
  asm volatile(
#if RVB_BF16
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#else
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
#endif
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// LKT = number of 16-key tiles (keys padded to 16*LKT)
template <int LKT>
__global__ void __launch_bounds__(256, (LKT <= 5) ? 4 : 1) bert_attn_kernel(const h16* __restrict__ qkv, h16* __restrict__ ctx, int L,
                                                        int heads) {
  RVB_PDL_PROLOGUE();
  constexpr int LP = LKT * 16;
  extern __shared__ __align__(16) uint8_t sm_raw[];
  h16* sQ = reinterpret_cast<h16*>(sm_raw);
  h16* sK = sQ + LP * PITCH;
  h16* sV = sK + LP * PITCH;
  const int head = blockIdx.x;
  const int row = blockIdx.y;
  const int H3 = heads * HD * 3;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  // stage Q, K, V head slices (zero-padded to LP rows)
  const h16* src = qkv + static_cast<long long>(row) * L * H3 + head * HD;
  // four 16-byte loads in flight per thread (the staging is the latency-bound part of this kernel)
  for (int i0 = threadIdx.x; i0 < LP * 8 * 3; i0 += 4 * blockDim.x) {
    uint4 val[4];
    h16* dst[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * blockDim.x;
      val[u] = make_uint4(0, 0, 0, 0);
      dst[u] = nullptr;
      if (i < LP * 8 * 3) {
        const int which = i / (LP * 8);
        const int rem = i - which * LP * 8;
        const int l = rem >> 3, v = rem & 7;
        if (l < L) val[u] = __ldg(reinterpret_cast<const uint4*>(src + static_cast<long long>(l) * H3 + which * heads * HD + v * 8));
        dst[u] = (which == 0 ? sQ : (which == 1 ? sK : sV)) + l * PITCH + v * 8;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (dst[u] != nullptr) *reinterpret_cast<uint4*>(dst[u]) = val[u];
  }
  __syncthreads();

  const int qtiles = (L + 15) / 16;
  for (int qt = warp; qt < qtiles; qt += nwarps) {
    // Q fragments for the 4 k-steps
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      const int r = qt * 16 + (lane & 15);
      const int c = ks * 16 + (lane >> 4) * 8;
      ldmatrix_x4(smem_u32(sQ + r * PITCH + c), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
    }
    float s[LKT * 2][4];
#pragma unroll
    for (int nt = 0; nt < LKT * 2; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f;
      const int r = nt * 8 + (lane & 7);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int c = half * 32 + (lane >> 3) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4(smem_u32(sK + r * PITCH + c), b0, b1, b2, b3);
        mma_h16_16816(s[nt], qa[half * 2][0], qa[half * 2][1], qa[half * 2][2], qa[half * 2][3], b0, b1);
        mma_h16_16816(s[nt], qa[half * 2 + 1][0], qa[half * 2 + 1][1], qa[half * 2 + 1][2], qa[half * 2 + 1][3], b2, b3);
      }
    }
    // softmax over keys (rows g = lane/4 and g+8), scale 1/sqrt(64)
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < LKT * 2; ++nt) {
      const int key = nt * 8 + (lane & 3) * 2;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool ok = (key + (j & 1)) < L;
        s[nt][j] = ok ? s[nt][j] * 0.125f : -INFINITY;
      }
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.0f, sum1 = 0.0f;
#pragma unroll
    for (int nt = 0; nt < LKT * 2; ++nt) {
      s[nt][0] = __expf(s[nt][0] - mx0);
      s[nt][1] = __expf(s[nt][1] - mx0);
      s[nt][2] = __expf(s[nt][2] - mx1);
      s[nt][3] = __expf(s[nt][3] - mx1);
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);

    // O = P V
    float o[8][4];
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.0f;
#pragma unroll
    for (int kt = 0; kt < LKT; ++kt) {
      const uint32_t a0 = pack_h2(s[2 * kt][0], s[2 * kt][1]);
      const uint32_t a1 = pack_h2(s[2 * kt][2], s[2 * kt][3]);
      const uint32_t a2 = pack_h2(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      const uint32_t a3 = pack_h2(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {  // pairs of 8-dim tiles
        const int r = kt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = (dp * 2 + (lane >> 4)) * 8;
        uint32_t b0, b1, b2, b3;
        ldmatrix_x4_trans(smem_u32(sV + r * PITCH + c), b0, b1, b2, b3);
        mma_h16_16816(o[dp * 2], a0, a1, a2, a3, b0, b1);
        mma_h16_16816(o[dp * 2 + 1], a0, a1, a2, a3, b2, b3);
      }
    }
    const float inv0 = 1.0f / sum0, inv1 = 1.0f / sum1;
    const int q0 = qt * 16 + (lane >> 2), q1 = q0 + 8;
    h16* out = ctx + static_cast<long long>(row) * L * heads * HD + head * HD + (lane & 3) * 2;
#pragma unroll
    for (int dt = 0; dt < 8; ++dt) {
      if (q0 < L)
        *reinterpret_cast<uint32_t*>(out + static_cast<long long>(q0) * heads * HD + dt * 8) =
            pack_h2(o[dt][0] * inv0, o[dt][1] * inv0);
      if (q1 < L)
        *reinterpret_cast<uint32_t*>(out + static_cast<long long>(q1) * heads * HD + dt * 8) =
            pack_h2(o[dt][2] * inv1, o[dt][3] * inv1);
    }
  }
}

template <int LKT>
void launch_bert_attn(const h16* qkv, h16* ctx, int R, int L, int heads, cudaStream_t s) {
  constexpr int LP = LKT * 16;
  const size_t smem = static_cast<size_t>(3) * LP * PITCH * sizeof(h16);
  static PerDeviceOnce attr_once;
  if (smem > 48 * 1024 && attr_once.first()) {
    RVB_CUDA(cudaFuncSetAttribute(bert_attn_kernel<LKT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  }
  const int qtiles = (L + 15) / 16;
  const int nwarps = qtiles < 8 ? qtiles : 8;
  dim3 grid(heads, R);
  launch_k(bert_attn_kernel<LKT>, dim3(grid), dim3(nwarps * 32), smem, s, qkv, ctx, L, heads);
  RVB_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------
// Visual_Ling_Attn cross attention: q [B*L,256] (shared by both modalities),
// kv [n_mod*B*16, 512] (k | v), ctx [n_mod*B*L, 256].  4 heads x 64, 16 keys.
// ---------------------------------------------------------------------------------------
constexpr int VH = 4, VK = 16, VP = HD + 4;   // pitch 68 floats: 16-byte aligned rows, odd number of 16 B chunks
__global__ void __launch_bounds__(256) vla_attn_kernel(const h16* __restrict__ q, const h16* __restrict__ kv,
                                                       h16* __restrict__ ctx, int B, int L, int q_shared) {
  RVB_PDL_PROLOGUE();
  __shared__ __align__(16) float sK[VH * VK * VP];
  __shared__ __align__(16) float sV[VH * VK * VP];
  const int b = blockIdx.x, mod = blockIdx.y;
  const h16* kvb = kv + (static_cast<long long>(mod) * B + b) * VK * 512;
  for (int i = threadIdx.x; i < VK * 64; i += blockDim.x) {   // 16-byte loads: 8 channels each
    const int j = i >> 6, c = (i & 63) * 8;
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(kvb + j * 512 + c));
    const int cc = c & 255, h = cc >> 6, d = cc & 63;
    float* dst = (c < 256 ? sK : sV) + (h * VK + j) * VP + d;
    const float2 t0 = unpack_h2(u.x), t1 = unpack_h2(u.y), t2 = unpack_h2(u.z), t3 = unpack_h2(u.w);
    *reinterpret_cast<float4*>(dst) = make_float4(t0.x, t0.y, t1.x, t1.y);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(t2.x, t2.y, t3.x, t3.y);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int h = lane >> 3, sub = lane & 7;
  // lane (h, sub): scores of keys sub and sub + 8 of head h; output dims [4 sub, 4 sub + 4) and
  // [32 + 4 sub, 32 + 4 sub + 4) -- every shared-memory access below is a conflict-free float4.
  // Query rows are split over gridDim.z CTAs (each stages the 16 KB of K/V again: cheaper than
  // leaving most SMs idle on this latency-bound step of the serial tail).
  const float4* k0 = reinterpret_cast<const float4*>(sK + (h * VK + sub) * VP);
  const float4* k1 = reinterpret_cast<const float4*>(sK + (h * VK + sub + 8) * VP);
  for (int l = blockIdx.z * nwarps + warp; l < L; l += nwarps * gridDim.z) {
    const uint4* qp = reinterpret_cast<const uint4*>(q + (static_cast<long long>(q_shared ? 0 : b) * L + l) * 256 + h * HD);
    uint4 qv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) qv[i] = __ldg(qp + i);
    float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float2 q0 = unpack_h2(qv[i].x), q1 = unpack_h2(qv[i].y), q2 = unpack_h2(qv[i].z), q3 = unpack_h2(qv[i].w);
      const float4 a0 = k0[2 * i], a1 = k0[2 * i + 1], b0 = k1[2 * i], b1 = k1[2 * i + 1];
      s0 = fmaf(q0.x, a0.x, s0); s0 = fmaf(q0.y, a0.y, s0); s0 = fmaf(q1.x, a0.z, s0); s0 = fmaf(q1.y, a0.w, s0);
      s0 = fmaf(q2.x, a1.x, s0); s0 = fmaf(q2.y, a1.y, s0); s0 = fmaf(q3.x, a1.z, s0); s0 = fmaf(q3.y, a1.w, s0);
      s1 = fmaf(q0.x, b0.x, s1); s1 = fmaf(q0.y, b0.y, s1); s1 = fmaf(q1.x, b0.z, s1); s1 = fmaf(q1.y, b0.w, s1);
      s1 = fmaf(q2.x, b1.x, s1); s1 = fmaf(q2.y, b1.y, s1); s1 = fmaf(q3.x, b1.z, s1); s1 = fmaf(q3.y, b1.w, s1);
    }
    s0 *= 0.125f; s1 *= 0.125f;
    float mx = fmaxf(s0, s1);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 4));
    float p0 = __expf(s0 - mx), p1 = __expf(s1 - mx);
    float sum = p0 + p1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    sum += __shfl_xor_sync(0xffffffffu, sum, 4);
    const float inv = 1.0f / sum;
    p0 *= inv; p1 *= inv;
    float o[8];
#pragma unroll
    for (int dd = 0; dd < 8; ++dd) o[dd] = 0.0f;
#pragma unroll
    for (int j = 0; j < VK; ++j) {
      const float pj = __shfl_sync(0xffffffffu, (j >> 3) ? p1 : p0, (h << 3) + (j & 7));
      const float4* vr = reinterpret_cast<const float4*>(sV + (h * VK + j) * VP);
      const float4 va = vr[sub], vb = vr[8 + sub];
      o[0] = fmaf(pj, va.x, o[0]); o[1] = fmaf(pj, va.y, o[1]); o[2] = fmaf(pj, va.z, o[2]); o[3] = fmaf(pj, va.w, o[3]);
      o[4] = fmaf(pj, vb.x, o[4]); o[5] = fmaf(pj, vb.y, o[5]); o[6] = fmaf(pj, vb.z, o[6]); o[7] = fmaf(pj, vb.w, o[7]);
    }
    h16* op = ctx + ((static_cast<long long>(mod) * B + b) * L + l) * 256 + h * HD;
    *reinterpret_cast<uint2*>(op + sub * 4) = make_uint2(pack_h2(o[0], o[1]), pack_h2(o[2], o[3]));
    *reinterpret_cast<uint2*>(op + 32 + sub * 4) = make_uint2(pack_h2(o[4], o[5]), pack_h2(o[6], o[7]));
  }
}

}  // namespace

void bert_self_attention(const h16* qkv, h16* ctx, int R, int L, int heads, cudaStream_t s) {
  RVB_CHECK(L >= 1 && L <= 256, "bert attention: 1 <= L <= 256 (INSTRUCTION_ENCODER.max_length is 200)");
  if (L <= 128 && heads * HD * 3 % 8 == 0 && use_tc_attention()) {   // tcgen05 / TMEM kernel (attention_tc.cu)
    bert_self_attention_tc(qkv, ctx, R, L, heads, s);
    return;
  }
  const int lkt = (L + 15) / 16;
  if (lkt <= 2) launch_bert_attn<2>(qkv, ctx, R, L, heads, s);
  else if (lkt <= 5) launch_bert_attn<5>(qkv, ctx, R, L, heads, s);
  else if (lkt <= 8) launch_bert_attn<8>(qkv, ctx, R, L, heads, s);
  else if (lkt <= 13) launch_bert_attn<13>(qkv, ctx, R, L, heads, s);
  else launch_bert_attn<16>(qkv, ctx, R, L, heads, s);
}

void vla_cross_attention(const h16* q, const h16* kv, h16* ctx, int B, int L, int n_mod, int q_shared,
                         cudaStream_t s) {
  const int zsplit = (B * n_mod >= 592) ? 1 : ((L + 31) / 32 > 4 ? 4 : (L + 31) / 32);
  dim3 grid(B, n_mod, zsplit);
  launch_k(vla_attn_kernel, dim3(grid), dim3(256), 0, s, q, kv, ctx, B, L, q_shared);
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
