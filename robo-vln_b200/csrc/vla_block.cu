// Fused cross-modal attention block of Visual_Ling_Attn (robo_vln_baselines/models/transformer/transformer.py:262-281
// with InterModuleAttnLayer :209-221, MultiHeadAttention / ScaledDotProductAttention :81-126 and
// PositionWiseFeedForward :38-43) as ONE tcgen05 / TMEM / TMA kernel.
//
// One CTA owns one (environment, modality) pair = one 128-row query tile (the L <= 128 instruction tokens of that
// environment against that modality's 16 visual cells) and runs the whole block without leaving the SM:
//
//   S   = Q0 . K'^T + c            tcgen05.mma  M=128 N=64  K=256   all 4 heads at once (K' = keys pulled through fc_q, below)
//   P   = softmax_16keys(S / 8)    registers <- TMEM (tcgen05.ld), un-normalised 16-bit P -> shared memory
//   O_h = P_h . V_h                tcgen05.mma  M=128 N=64  K=16    V consumed MN-major straight from its TMA box
//   ctx = O / rowsum               TMEM -> 16-bit shared-memory A operand (K-major, 128B swizzle)
//   A   = ctx . Wo^T               tcgen05.mma  M=128 N=256 K=256   Wo streamed by TMA through a 5-slot ring
//   X   = LN1(Q0 + A + bo)         TMEM -> registers -> 16-bit shared-memory A operand (overwrites ctx)
//   H_c = relu(X . W1_c^T + b1_c)  8 chunks of 128 hidden units, TMEM accumulators double-buffered,
//   Y  += H_c . W2_c^T               so that chunk c's activation epilogue runs under the MMAs of chunks c+1 / c-1
//   Y   = LN2(X + Y + b2)          TMEM -> registers -> shared memory (in place of X)
//   out = mean_{l<L} Y[l, :]       AdaptiveAvgPool1d over the tokens (seq2seq_highlevel_cma.py:200-210)
//
// Nothing but the pooled [256] vector is written to global memory; the [rows, 1024] hidden activation, the
// attention context and both LayerNorm outputs of the unfused path never exist in HBM.
//
// The query projection is folded into the KEY side: S_h = (Q0 Wq_h^T + bq_h) K_h^T = Q0 (K_h Wq_h)^T + K_h bq_h, so the
// 16 keys of a cell grid are pulled through fc_q (16 rows) instead of the L queries through it (L rows): the producing
// GEMM ("kvx", engine.cu) emits, per visual cell, [K'(head 0..3) x 256 | c(4, padded to 8) | V(256)] with the composed
// weights (Wq_h^T Wk_h, ...) prepared once in fp32 (weight_prep.py).  Q0 = LN0(relu(ins_fc(bert))) + PE comes from the
// LayerNorm-epilogue GEMM (gemm_tc.cu) and is both the A operand of S and the residual of LN1, read from shared memory.
//
// Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer, warps 2..9 = epilogue:
// two warps per TMEM lane quadrant, each owning one half of the columns of every stage (row statistics of the two
// halves are exchanged through shared memory with a 64-thread named barrier).
#include "common.cuh"
#include "rvb.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace rvb {

namespace {

constexpr int VB_THREADS = 320;
constexpr int VB_EPI_WARPS = 8;
constexpr int VB_NS = 5;                         // weight-ring slots
constexpr int VB_SLOT = 16384;                   // bytes per slot: one [128 rows x 64 k] 16-bit block
constexpr int VB_SUB = 16384;                    // one [128 rows x 64 cols] K-major sub-tile of an A operand
constexpr int VB_OFF_A = 0;                      // bufA: Q0 (4 sub-tiles), later the two H chunk buffers (2 x 2 sub-tiles)
constexpr int VB_OFF_B = 65536;                  // bufB: (P, V) -> ctx -> X -> Y (4 sub-tiles)
constexpr int VB_OFF_P = VB_OFF_B;               // P: [128 x 64] un-normalised probabilities (col = head*16 + key); dead once P.V
constexpr int VB_OFF_V = VB_OFF_B + 16384;       // V: 4 heads x [16 keys x 64 dims];                                  has retired
constexpr int VB_OFF_RING = 131072;
constexpr int VB_OFF_MISC = VB_OFF_RING + VB_NS * VB_SLOT;
constexpr int VB_MISC_BYTES = 4096;              // barriers, c[64], LayerNorm partials [2][128] float2
constexpr int VB_OFF_PAR = VB_OFF_MISC + VB_MISC_BYTES;   // fp32 bo | ln1g | ln1b | b2 | ln2g | ln2b (256 each) | b1 (1024)
constexpr int VB_PAR_FLOATS = 6 * 256 + 1024;
constexpr int VB_SMEM = VB_OFF_PAR + VB_PAR_FLOATS * 4 + 1024 /*align slack*/;
static_assert(VB_SMEM <= 232448, "vla_block: shared memory over the 227 KB limit");

constexpr int TM_Y = 0;       // TMEM columns: O (P.V) and later the fc2 accumulator Y
constexpr int TM_H = 256;     // S, then the fc_o accumulator (256 cols), then the two fc1 chunk accumulators (2 x 128)

struct VlaBlockParams {
  int B, L;
  int q_shared;                 // 1: one instruction for every environment (Q0 has L rows)
  const h16* kvx;               // [2*B*16, kvx_pitch]: K' | c | V per visual cell (c read directly)
  long long kvx_pitch;
  const float *bo, *b1, *b2, *ln1g, *ln1b, *ln2g, *ln2b;
  float eps;
  h16* out;                     // pooled [B, out_pitch], modality m at column m*256
  long long out_pitch;
  h16* y_tokens;                // parity tests only: token-level output [2, B, L, 256]; null in production
  int rotate;                   // 1: every CTA walks the weight blocks in its own rotated order (see below)
  long long* times;             // diagnostics (ROBOVLN_VLA_TIMES): [CTA][64] SM-clock stamps; null in production
};

// diagnostics: slot i of this CTA's timeline (MMA thread: 0..31, first epilogue thread: 32..63)
#define VB_STAMP(i)                                                                                   \
  do {                                                                                                \
    if (p.times != nullptr) p.times[(static_cast<long long>(blockIdx.y) * gridDim.x + blockIdx.x) * 64 + (i)] = clock64(); \
  } while (0)

RVB_DEVICE void named_bar(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

RVB_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// 128 fp32 accumulator columns of this thread's TMEM lane -> registers
RVB_DEVICE void tmem_ld_128(uint32_t taddr, float (&f)[128]) {
  uint32_t u[4][32];
#pragma unroll
  for (int i = 0; i < 4; ++i) tmem_ld_32x32(taddr + 32 * i, u[i]);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 32; ++j) f[32 * i + j] = __uint_as_float(u[i][j]);
}

// 8 consecutive 16-bit values (one 16-byte chunk of a swizzled K-major row) -> fp32
RVB_DEVICE void add_chunk(float* f, const uint4 r) {
  float2 t;
  t = unpack_h2(r.x); f[0] += t.x; f[1] += t.y;
  t = unpack_h2(r.y); f[2] += t.x; f[3] += t.y;
  t = unpack_h2(r.z); f[4] += t.x; f[5] += t.y;
  t = unpack_h2(r.w); f[6] += t.x; f[7] += t.y;
}
RVB_DEVICE uint4 pack_chunk(const float* f) {
  uint4 q;
  q.x = pack_h2(f[0], f[1]); q.y = pack_h2(f[2], f[3]); q.z = pack_h2(f[4], f[5]); q.w = pack_h2(f[6], f[7]);
  return q;
}

// The FFN's MMA issue order: fc1(0), then for c = 1..7 { fc1(c), fc2(c-1) }, then fc2(7): chunk c's activation
// epilogue overlaps fc1(c+1) and fc2(c-1).  step 0..15 -> (is_fc2, chunk)
RVB_DEVICE void ffn_step(int step, int& is_fc2, int& chunk) {
  if (step == 0) { is_fc2 = 0; chunk = 0; }
  else if (step == 15) { is_fc2 = 1; chunk = 7; }
  else { is_fc2 = (step & 1) ? 0 : 1; chunk = (step & 1) ? (step + 1) >> 1 : (step >> 1) - 1; }
}

// LayerNorm epilogue over the 256 fp32 accumulator columns of this thread's row, of which this thread owns 128
// (tcol = first of them); the partner warp of the same lane quadrant owns the other 128.
//   v = acc + bias + residual (16-bit, swizzled K-major sub-tiles res0 / res0 + VB_SUB)        pass 1, written back to TMEM
//   y = (v - mean) * rstd * gamma + beta -> 16-bit, same layout, dst0 / dst0 + VB_SUB            pass 2
// Two passes over TMEM keep the live register set small, so the compiler can keep many shared-memory loads in flight
// (holding all 128 values in registers measured 8.5k cycles per LayerNorm, latency bound).  Parameters come from the
// shared-memory copy made at kernel start: with 218 KB of shared memory in use the L1 has ~10 KB left and global
// parameter loads thrash it.
template <class OnChunk>
RVB_DEVICE void ln_epilogue(uint32_t tcol, const uint8_t* res0, uint8_t* dst0, const float* bias, const float* gamma,
                            const float* beta, float2* xchg, int half, int row, int quad, float eps, OnChunk&& on_chunk,
                            long long* stamps = nullptr) {
  const int sw = row & 7;
  float s = 0.0f, q = 0.0f;
  // TMEM loads are double-buffered in registers: chunk ch + 1 is in flight while chunk ch is being processed
  // (tcgen05.wait::ld has no per-load granularity, so the wait for ch + 1 sits after the arithmetic of ch)
  uint32_t u[2][32];
  tmem_ld_32x32(tcol, u[0]);
  tmem_ld_wait();
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (ch < 3) tmem_ld_32x32(tcol + (ch + 1) * 32, u[(ch + 1) & 1]);
    uint4 r[4];
    float4 bb[8];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      r[j] = *reinterpret_cast<const uint4*>(res0 + (ch >> 1) * VB_SUB + row * 128 + ((((ch & 1) * 4 + j) ^ sw) << 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) bb[j] = *reinterpret_cast<const float4*>(bias + ch * 32 + j * 4);
    float f[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      f[4 * j] = __uint_as_float(u[ch & 1][4 * j]) + bb[j].x; f[4 * j + 1] = __uint_as_float(u[ch & 1][4 * j + 1]) + bb[j].y;
      f[4 * j + 2] = __uint_as_float(u[ch & 1][4 * j + 2]) + bb[j].z; f[4 * j + 3] = __uint_as_float(u[ch & 1][4 * j + 3]) + bb[j].w;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) add_chunk(&f[8 * j], r[j]);
    uint32_t w[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      s += f[j];
      q = fmaf(f[j], f[j], q);
      w[j] = __float_as_uint(f[j]);
    }
    tmem_st_32x32(tcol + ch * 32, w);
    if (ch < 3) tmem_ld_wait();
  }
  tmem_st_wait();
  if (stamps != nullptr) stamps[0] = clock64();
  xchg[half * 128 + row] = make_float2(s, q);
  named_bar(1 + quad, 64);
  const float2 p0 = xchg[row], p1 = xchg[128 + row];     // fixed order: both halves compute identical totals
  named_bar(1 + quad, 64);                               // the slots may be rewritten by the next LayerNorm
  if (stamps != nullptr) stamps[1] = clock64();
  const float mean = (p0.x + p1.x) * (1.0f / 256.0f);
  const float var = fmaxf((p0.y + p1.y) * (1.0f / 256.0f) - mean * mean, 0.0f);
  const float a = rsqrtf(var + eps);
  const float b = -mean * a;
  tmem_ld_32x32(tcol, u[0]);
  tmem_ld_wait();
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (ch < 3) tmem_ld_32x32(tcol + (ch + 1) * 32, u[(ch + 1) & 1]);
    float4 gg[8], ee[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      gg[j] = *reinterpret_cast<const float4*>(gamma + ch * 32 + j * 4);
      ee[j] = *reinterpret_cast<const float4*>(beta + ch * 32 + j * 4);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t* uc = &u[ch & 1][8 * j];
      float g[8];
      g[0] = fmaf(fmaf(__uint_as_float(uc[0]), a, b), gg[2 * j].x, ee[2 * j].x);
      g[1] = fmaf(fmaf(__uint_as_float(uc[1]), a, b), gg[2 * j].y, ee[2 * j].y);
      g[2] = fmaf(fmaf(__uint_as_float(uc[2]), a, b), gg[2 * j].z, ee[2 * j].z);
      g[3] = fmaf(fmaf(__uint_as_float(uc[3]), a, b), gg[2 * j].w, ee[2 * j].w);
      g[4] = fmaf(fmaf(__uint_as_float(uc[4]), a, b), gg[2 * j + 1].x, ee[2 * j + 1].x);
      g[5] = fmaf(fmaf(__uint_as_float(uc[5]), a, b), gg[2 * j + 1].y, ee[2 * j + 1].y);
      g[6] = fmaf(fmaf(__uint_as_float(uc[6]), a, b), gg[2 * j + 1].z, ee[2 * j + 1].z);
      g[7] = fmaf(fmaf(__uint_as_float(uc[7]), a, b), gg[2 * j + 1].w, ee[2 * j + 1].w);
      const uint4 yq = pack_chunk(g);
      *reinterpret_cast<uint4*>(dst0 + (ch >> 1) * VB_SUB + row * 128 + ((((ch & 1) * 4 + j) ^ sw) << 4)) = yq;
      on_chunk(ch * 32 + j * 8, yq);
    }
    if (ch < 3) tmem_ld_wait();
  }
}

__global__ void __launch_bounds__(VB_THREADS, 1)
vla_block_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmKp,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmWo,
                 const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const VlaBlockParams p) {
  extern __shared__ uint8_t vb_raw[];
  uint8_t* smem = vb_raw + ((1024u - (smem_u32(vb_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* bufA = smem + VB_OFF_A;
  uint8_t* bufB = smem + VB_OFF_B;
  uint8_t* sP = smem + VB_OFF_P;
  uint8_t* sV = smem + VB_OFF_V;
  uint8_t* ring = smem + VB_OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + VB_OFF_MISC);
  uint64_t* full_bar = bars;                 // [VB_NS]
  uint64_t* empty_bar = bars + VB_NS;        // [VB_NS]
  uint64_t* one = bars + 2 * VB_NS;          // single-use: 0 in, 1 s_done, 2 p_ready, 3 o_done, 4 ctx_ready, 5 fco_done, 6 x_ready, 7 y_done
  uint64_t* hfull = one + 8;                 // [2] fc1 chunk accumulator complete        (MMA -> epilogue)
  uint64_t* hempty = hfull + 2;              // [2] fc1 chunk accumulator drained         (epilogue -> MMA)
  uint64_t* sfull = hempty + 2;              // [2] H chunk written to shared memory      (epilogue -> MMA)
  uint64_t* sempty = sfull + 2;              // [2] fc2 finished reading the H chunk      (MMA -> epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sempty + 2);
  float* s_c = reinterpret_cast<float*>(smem + VB_OFF_MISC + 256);          // [64] c[key*4 + head]
  float2* s_ln = reinterpret_cast<float2*>(smem + VB_OFF_MISC + 512);       // [2][128]
  float* s_par = reinterpret_cast<float*>(smem + VB_OFF_PAR);               // bo | ln1g | ln1b | b2 | ln2g | ln2b | b1

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform: the role branches are uniform branches
  const int lane = threadIdx.x & 31;
  const int env = blockIdx.x, mod = blockIdx.y;
  const int cell_row0 = (mod * p.B + env) * 16;          // first of this tile's 16 visual-cell rows in kvx
  const int q_row0 = p.q_shared ? 0 : env * p.L;
  // Every CTA streams the SAME 1.1 MB of weights; walking them in lock step makes all SMs ask the same L2 lines at
  // the same moment (measured: ~0.75 us per 16 KB block, 4x the tensor-pipe time).  The sum over k blocks and over the
  // FFN's hidden chunks is order independent, so CTA i starts at chunk (i mod 8) and k block ((i / 8) mod 4): at any
  // instant the CTAs are spread over 32 different blocks.  (Deterministic: the order is a function of the CTA index.)
  const int cta_id = p.rotate ? static_cast<int>(blockIdx.y * gridDim.x + blockIdx.x) : 0;
  const int rc = cta_id & 7, rk = (cta_id >> 3) & 3;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmKp); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int i = 0; i < VB_NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(&one[0], 1); mbar_init(&one[1], 1); mbar_init(&one[2], VB_EPI_WARPS); mbar_init(&one[3], 1);
    mbar_init(&one[4], VB_EPI_WARPS); mbar_init(&one[5], 1); mbar_init(&one[6], VB_EPI_WARPS); mbar_init(&one[7], 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&hfull[i], 1); mbar_init(&hempty[i], VB_EPI_WARPS);
      mbar_init(&sfull[i], VB_EPI_WARPS); mbar_init(&sempty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  RVB_PDL_PROLOGUE();

  if (warp == 0) {
    // =============================== TMA producer ===============================
    {   // the whole warp walks this block convergently (operands in uniform registers); lane 0 issues
      const bool issuer = (lane == 0);
      // resident operands: Q0 tile (4 K blocks) and the 4 per-head V tiles
      if (issuer) mbar_arrive_expect_tx(&one[0], 4 * VB_SUB + 4 * 2048);
      for (int kb = 0; kb < 4; ++kb) if (issuer) tma_load_2d(bufA + kb * VB_SUB, &tmQ0, &one[0], kb * 64, q_row0);
      for (int h = 0; h < 4; ++h) if (issuer) tma_load_2d(sV + h * 2048, &tmV, &one[0], 1032 + h * 64, cell_row0);
      int slot = 0;
      uint32_t phase = 0;
      auto next = [&]() { if (++slot == VB_NS) { slot = 0; phase ^= 1; } };
      // S: K' blocks [64 (key, head) rows x 64 k]
      for (int kb = 0; kb < 4; ++kb) {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (issuer) mbar_arrive_expect_tx(&full_bar[slot], 8192);
        if (issuer) tma_load_3d(ring + slot * VB_SLOT, &tmKp, &full_bar[slot], kb * 64, 0, cell_row0);
        next();
      }
      // fc_o: Wo [256 x 256] as (k block, n half) units of [128 x 64]
      for (int kb = 0; kb < 4; ++kb)
        for (int nh = 0; nh < 2; ++nh) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          if (issuer) mbar_arrive_expect_tx(&full_bar[slot], VB_SLOT);
          if (issuer) tma_load_2d(ring + slot * VB_SLOT, &tmWo, &full_bar[slot], ((kb + rk) & 3) * 64, ((nh + rc) & 1) * 128);
          next();
        }
      // FFN in MMA issue order
      for (int step = 0; step < 16; ++step) {
        int is_fc2, c;
        ffn_step(step, is_fc2, c);
        for (int u = 0; u < 4; ++u) {
          mbar_wait(&empty_bar[slot], phase ^ 1);
          if (issuer) mbar_arrive_expect_tx(&full_bar[slot], VB_SLOT);
          const int cc = (c + rc) & 7;       // the hidden chunk this CTA processes at sequence position c
          if (issuer) {
            if (!is_fc2) tma_load_2d(ring + slot * VB_SLOT, &tmW1, &full_bar[slot], ((u + rk) & 3) * 64, cc * 128);    // W1[cc*128.., k block]
            else tma_load_2d(ring + slot * VB_SLOT, &tmW2, &full_bar[slot], cc * 128 + (((u >> 1) + rk) & 1) * 64,
                             (((u & 1) + (rk >> 1)) & 1) * 128);                                                        // W2[n half, k = cc*128 + kb2*64]
          }
          next();
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    {   // the whole warp walks this block convergently (operands in uniform registers); lane 0 issues
      const bool issuer = (lane == 0);
      int slot = 0;
      uint32_t phase = 0;
      auto next = [&]() { if (++slot == VB_NS) { slot = 0; phase ^= 1; } };
      // one ring unit: 4 k-steps of 16 against A sub-tile `a_sub`, accumulating into TMEM column d_col
      auto unit = [&](const uint8_t* a_sub, uint32_t d_col, uint32_t idesc, bool first) {
        mbar_wait(&full_bar[slot], phase);
        tc_fence_after();
        const uint64_t adesc = umma_desc_sw128(smem_u32(a_sub));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + slot * VB_SLOT));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          if (issuer) umma_f16kind(tmem_base + d_col, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                       static_cast<uint32_t>(!(first && k == 0)));
        if (issuer) umma_commit(&empty_bar[slot]);
        next();
      };
      constexpr uint32_t idesc64 = umma_idesc_h16(128, 64);
      constexpr uint32_t idesc128 = umma_idesc_h16(128, 128);
      if (issuer) VB_STAMP(0);
      mbar_wait(&one[0], 0);
      tc_fence_after();
      if (issuer) VB_STAMP(1);
      // ---- S[128, 64] = Q0 . K'^T
      for (int kb = 0; kb < 4; ++kb) unit(bufA + kb * VB_SUB, TM_H, idesc64, kb == 0);
      if (issuer) umma_commit(&one[1]);
      if (issuer) VB_STAMP(2);
      // ---- O_h[128, 64] = P_h . V_h   (V MN-major: [key, dim] rows of 128 B)
      mbar_wait(&one[2], 0);
      tc_fence_after();
      if (issuer) VB_STAMP(3);
      {
        constexpr uint32_t idesc_mn = umma_idesc_h16(128, 64) | (1u << 16);
        const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
        for (int h = 0; h < 4; ++h)
          if (issuer) umma_f16kind(tmem_base + TM_Y + h * 64, pdesc + static_cast<uint64_t>(h * 2), umma_desc_sw128(smem_u32(sV + h * 2048)),
                       idesc_mn, 0u);
        if (issuer) umma_commit(&one[3]);
      }
      // ---- fc_o: A[128, 256] = ctx . Wo^T
      mbar_wait(&one[4], 0);
      tc_fence_after();
      if (issuer) VB_STAMP(4);
      for (int kb = 0; kb < 4; ++kb)
        for (int nh = 0; nh < 2; ++nh) unit(bufB + ((kb + rk) & 3) * VB_SUB, TM_H + ((nh + rc) & 1) * 128, idesc128, kb == 0);
      if (issuer) umma_commit(&one[5]);
      if (issuer) VB_STAMP(5);
      // ---- FFN
      mbar_wait(&one[6], 0);
      tc_fence_after();
      if (issuer) VB_STAMP(6);
      for (int step = 0; step < 16; ++step) {
        if (issuer) VB_STAMP(8 + step);
        int is_fc2, c;
        ffn_step(step, is_fc2, c);
        const int b = c & 1;
        if (!is_fc2) {
          if (c >= 2) {   // the epilogue has drained this accumulator (chunk c - 2)
            mbar_wait(&hempty[b], static_cast<uint32_t>(((c >> 1) - 1) & 1));
            tc_fence_after();
          }
          for (int kb = 0; kb < 4; ++kb) unit(bufB + ((kb + rk) & 3) * VB_SUB, TM_H + b * 128, idesc128, kb == 0);
          if (issuer) umma_commit(&hfull[b]);
        } else {
          mbar_wait(&sfull[b], static_cast<uint32_t>((c >> 1) & 1));
          tc_fence_after();
          for (int u = 0; u < 4; ++u)
            unit(bufA + b * 2 * VB_SUB + (((u >> 1) + rk) & 1) * VB_SUB, TM_Y + (((u & 1) + (rk >> 1)) & 1) * 128, idesc128,
                 c == 0 && (u >> 1) == 0);
          if (issuer) umma_commit(&sempty[b]);
        }
      }
      if (issuer) umma_commit(&one[7]);
      if (issuer) VB_STAMP(7);
    }
  } else {
    // =============================== epilogue (warps 2..9) ===============================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int sw = row & 7;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int et = threadIdx.x - 64;      // 0..255
    // c[key*4 + head] = K_h[key] . bq_h  (kvx columns 1024..1027 of each visual-cell row)
    if (et < 64) s_c[et] = from_h16(p.kvx[static_cast<long long>(cell_row0 + (et >> 2)) * p.kvx_pitch + 1024 + (et & 3)]);
    {   // epilogue parameters -> shared memory (constant weights: 10 KB, read ~100 times per thread below)
      const float* srcs[6] = {p.bo, p.ln1g, p.ln1b, p.b2, p.ln2g, p.ln2b};
#pragma unroll
      for (int i = 0; i < 6; ++i) s_par[i * 256 + et] = __ldg(srcs[i] + et);
#pragma unroll
      for (int i = 0; i < 4; ++i) s_par[6 * 256 + i * 256 + et] = __ldg(p.b1 + i * 256 + et);
    }
    named_bar(5, VB_EPI_WARPS * 32);

    // ---- softmax over the 16 keys of this thread's two heads (2*half, 2*half + 1); S column = key*4 + head
    float inv_sum[2];
    {
      if (et == 0) VB_STAMP(32);
      mbar_wait(&one[1], 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(33);
      float s[64];
      {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(trow + TM_H, v0);
        tmem_ld_32x32(trow + TM_H + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { s[j] = __uint_as_float(v0[j]); s[32 + j] = __uint_as_float(v1[j]); }
      }
      const float scale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        float x[16];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          x[k] = (half == 0 ? s[k * 4 + hh] : s[k * 4 + 2 + hh]) + s_c[k * 4 + h];
          mx = fmaxf(mx, x[k]);
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          x[k] = exp2f((x[k] - mx) * scale);
          sum += x[k];
        }
        inv_sum[hh] = 1.0f / sum;
        // P row: column h*16 + key  -> 16-byte chunks 2h, 2h + 1
        *reinterpret_cast<uint4*>(sP + row * 128 + (((2 * h) ^ sw) << 4)) = pack_chunk(&x[0]);
        *reinterpret_cast<uint4*>(sP + row * 128 + (((2 * h + 1) ^ sw) << 4)) = pack_chunk(&x[8]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&one[2]);
    }

    // ---- ctx = O / rowsum for heads 2*half, 2*half + 1 -> bufB sub-tiles (K block = head)
    {
      if (et == 0) VB_STAMP(34);
      mbar_wait(&one[3], 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(35);
      float f[128];
      tmem_ld_128(trow + TM_Y + half * 128, f);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint8_t* dst = bufB + (2 * half + hh) * VB_SUB + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float g[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = f[hh * 64 + j * 8 + e] * inv_sum[hh];
          *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) = pack_chunk(g);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&one[4]);
    }

    // ---- X = LN1(Q0 + A + bo): this thread owns columns half*128 .. +127
    {
      if (et == 0) VB_STAMP(36);
      mbar_wait(&one[0], 0);     // (long complete) this thread reads the TMA-written Q0 tile below
      mbar_wait(&one[5], 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(37);
      const int n0 = half * 128;
      ln_epilogue(trow + TM_H + n0, bufA + 2 * half * VB_SUB, bufB + 2 * half * VB_SUB, s_par + n0, s_par + 256 + n0,
                  s_par + 512 + n0, s_ln, half, row, quad, p.eps, [](int, const uint4&) {});
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&one[6]);
    }

    // ---- H_c = relu(fc1 chunk + b1): this thread owns 64 of the chunk's 128 hidden units = sub-tile `half`
#pragma unroll 1
    for (int c = 0; c < 8; ++c) {
      const int b = c & 1;
      if (et == 0) VB_STAMP(40 + 2 * c);
      mbar_wait(&hfull[b], static_cast<uint32_t>((c >> 1) & 1));
      tc_fence_after();
      if (et == 0) VB_STAMP(41 + 2 * c);
      uint32_t r0[32], r1[32];
      tmem_ld_32x32(trow + TM_H + b * 128 + half * 64, r0);
      tmem_ld_32x32(trow + TM_H + b * 128 + half * 64 + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&hempty[b]);           // accumulator drained: fc1(c + 2) may overwrite it
      if (c >= 2) mbar_wait(&sempty[b], static_cast<uint32_t>(((c >> 1) - 1) & 1));   // fc2(c - 2) has read this H buffer
      uint8_t* dst = bufA + (b * 2 + half) * VB_SUB + row * 128;
      const float* bias = s_par + 6 * 256 + ((c + rc) & 7) * 128 + half * 64;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + j * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + j * 8 + 4);
        const uint32_t* src = (j < 4) ? &r0[j * 8] : &r1[(j - 4) * 8];
        float g[8];
        g[0] = fmaxf(__uint_as_float(src[0]) + b0.x, 0.0f); g[1] = fmaxf(__uint_as_float(src[1]) + b0.y, 0.0f);
        g[2] = fmaxf(__uint_as_float(src[2]) + b0.z, 0.0f); g[3] = fmaxf(__uint_as_float(src[3]) + b0.w, 0.0f);
        g[4] = fmaxf(__uint_as_float(src[4]) + b1.x, 0.0f); g[5] = fmaxf(__uint_as_float(src[5]) + b1.y, 0.0f);
        g[6] = fmaxf(__uint_as_float(src[6]) + b1.z, 0.0f); g[7] = fmaxf(__uint_as_float(src[7]) + b1.w, 0.0f);
        *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) = pack_chunk(g);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[b]);
    }

    // ---- Y = LN2(X + Y + b2), written in place of X; then the token mean
    {
      if (et == 0) VB_STAMP(38);
      mbar_wait(&one[7], 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(39);
      const int n0 = half * 128;
      h16* ytok = (p.y_tokens != nullptr && row < p.L) ? p.y_tokens + ((static_cast<long long>(mod) * p.B + env) * p.L + row) * 256 + n0 : nullptr;
      ln_epilogue(trow + TM_Y + n0, bufB + 2 * half * VB_SUB, bufB + 2 * half * VB_SUB, s_par + 768 + n0, s_par + 1024 + n0,
                  s_par + 1280 + n0, s_ln, half, row, quad, p.eps, [ytok](int col, const uint4& yq) {
                    if (ytok != nullptr) *reinterpret_cast<uint4*>(ytok + col) = yq;
                  });
      named_bar(5, VB_EPI_WARPS * 32);
      // column `et` over the valid token rows (16-bit Y, fp32 sum, fixed order)
      const uint8_t* colp = bufB + (et >> 6) * VB_SUB + (et & 7) * 2;
      const int chunk = (et & 63) >> 3;
      float acc4[4] = {0.0f, 0.0f, 0.0f, 0.0f};     // rows r = 4i + k go to accumulator k: independent chains, fixed order
      int r = 0;
      for (; r + 8 <= p.L; r += 8) {
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = from_h16(*reinterpret_cast<const h16*>(colp + (r + k) * 128 + ((chunk ^ k) << 4)));   // (r + k) & 7 == k
#pragma unroll
        for (int k = 0; k < 8; ++k) acc4[k & 3] += x[k];
      }
      // tail: r is a multiple of 8 here, so row r + k has k = (r + k) & 7 and goes to accumulator k & 3 -- static indices
      // (indexing acc4 with the run-time row number had put the four accumulators in local memory)
#pragma unroll
      for (int k = 0; k < 7; ++k)
        if (r + k < p.L) acc4[k & 3] += from_h16(*reinterpret_cast<const h16*>(colp + (r + k) * 128 + ((chunk ^ k) << 4)));
      const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      p.out[static_cast<long long>(env) * p.out_pitch + mod * 256 + et] = to_h16(acc / static_cast<float>(p.L));
      if (et == 0) VB_STAMP(56);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}


// =====================================================================================================================
// CTA-pair form (cta_group::2): the two modalities of ONE environment share a 2-CTA cluster (rank 0 = RGB tile,
// rank 1 = depth tile).  Every weight GEMM is one M=256 UMMA across the pair -- each CTA stages its own 128 query rows
// (A) and HALF of every weight block (B) -- which halves, per SM, both the L2->SM weight traffic and the shared-memory
// bandwidth the MMA needs (the 1-CTA form above is bound by exactly that: 32 KB of operand reads + 16 KB of TMA writes
// per 256 tensor-pipe cycles), and lets the FFN run with 256-wide hidden chunks (N = 256 MMAs).
//   S  : [Q0; Q0] . [K'_rgb ; K'_depth]^T   M=256 N=128: each CTA stages ITS OWN K' as its half of B and reads back ITS
//        64 columns of the result (the other 64 are the queries against the other modality's keys: ignored)
//   O_h: P . [V_rgb,h | V_depth,h]          M=256 N=128 K=16, same trick with the MN-major V tiles
//   fc_o, fc1 (4 chunks of 256), fc2        M=256 N=256, B half = weight rows [rank*128, +128) of the block
// The leader (rank 0) issues every MMA; completion is broadcast to both CTAs (tcgen05.commit multicast); the epilogue
// warps of both CTAs report to the leader's barriers through the cluster shared-memory window.
// =====================================================================================================================
RVB_DEVICE void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* map, uint32_t mbar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

constexpr int VP_NS = 5;

__global__ void __launch_bounds__(VB_THREADS, 1)
vla_pair_kernel(const __grid_constant__ CUtensorMap tmQ0, const __grid_constant__ CUtensorMap tmKp,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmWo,
                const __grid_constant__ CUtensorMap tmW1, const __grid_constant__ CUtensorMap tmW2, const VlaBlockParams p) {
  extern __shared__ uint8_t vb_raw[];
  uint8_t* smem = vb_raw + ((1024u - (smem_u32(vb_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS, not generic LD / ST)
  uint8_t* bufA = smem + VB_OFF_A;
  uint8_t* bufB = smem + VB_OFF_B;
  uint8_t* sP = smem + VB_OFF_P;
  uint8_t* sV = smem + VB_OFF_V;
  uint8_t* ring = smem + VB_OFF_RING;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + VB_OFF_MISC);
  uint64_t* full_bar = bars;                 // [VP_NS]  leader only: bytes of both CTAs' halves
  uint64_t* empty_bar = bars + VP_NS;        // [VP_NS]  both CTAs (multicast commit)
  uint64_t* in_bar = bars + 2 * VP_NS;       // local: Q0 + V landed
  uint64_t* in_both = in_bar + 1;            // leader: both CTAs' inputs landed (2 arrivals)
  uint64_t* s_done = in_bar + 2;             // multicast commits ...
  uint64_t* o_done = in_bar + 3;
  uint64_t* fco_done = in_bar + 4;
  uint64_t* y_done = in_bar + 5;
  uint64_t* hfull = in_bar + 6;
  uint64_t* sempty = in_bar + 7;
  uint64_t* p_ready = in_bar + 8;            // leader: arrivals of all 16 epilogue warps of the pair ...
  uint64_t* ctx_ready = in_bar + 9;
  uint64_t* x_ready = in_bar + 10;
  uint64_t* hempty = in_bar + 11;
  uint64_t* sfull = in_bar + 12;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(in_bar + 13);
  float* s_c = reinterpret_cast<float*>(smem + VB_OFF_MISC + 256);
  float2* s_ln = reinterpret_cast<float2*>(smem + VB_OFF_MISC + 512);
  float* s_par = reinterpret_cast<float*>(smem + VB_OFF_PAR);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);   // provably warp-uniform: the role branches are uniform branches
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = RGB tile (leader), 1 = depth tile
  const int env = static_cast<int>(blockIdx.x >> 1), mod = static_cast<int>(rank);
  const int cell_row0 = (mod * p.B + env) * 16;
  const int q_row0 = p.q_shared ? 0 : env * p.L;
  auto leader_addr = [](uint64_t* bar) { return mapa_u32(smem_u32(bar), 0); };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ0); tma_prefetch_desc(&tmKp); tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmWo); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    for (int i = 0; i < VP_NS; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(in_bar, 1); mbar_init(in_both, 2);
    mbar_init(s_done, 1); mbar_init(o_done, 1); mbar_init(fco_done, 1); mbar_init(y_done, 1); mbar_init(hfull, 1); mbar_init(sempty, 1);
    mbar_init(p_ready, 2 * VB_EPI_WARPS); mbar_init(ctx_ready, 2 * VB_EPI_WARPS); mbar_init(x_ready, 2 * VB_EPI_WARPS);
    mbar_init(hempty, 2 * VB_EPI_WARPS); mbar_init(sfull, 2 * VB_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer's barriers exist before any remote arrive / multicast commit / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  RVB_PDL_PROLOGUE();

  if (warp == 0) {
    // =============================== TMA producer (both CTAs) ===============================
    {   // the whole warp walks this block convergently (operands in uniform registers); lane 0 issues
      const bool issuer = (lane == 0);
      if (issuer) mbar_arrive_expect_tx(in_bar, 4 * VB_SUB + 4 * 2048);
      for (int kb = 0; kb < 4; ++kb) if (issuer) tma_load_2d(bufA + kb * VB_SUB, &tmQ0, in_bar, kb * 64, q_row0);
      for (int h = 0; h < 4; ++h) if (issuer) tma_load_2d(sV + h * 2048, &tmV, in_bar, 1032 + h * 64, cell_row0);
      int slot = 0;
      uint32_t phase = 0;
      auto next = [&]() { if (++slot == VP_NS) { slot = 0; phase ^= 1; } };
      // every unit: wait until the pair's MMAs have retired the slot, then stage THIS CTA's half; the bytes of both
      // halves are credited to the leader's full barrier, which the leader arms
      auto begin_unit = [&](uint32_t bytes_per_cta) -> uint32_t {
        mbar_wait(&empty_bar[slot], phase ^ 1);
        if (rank == 0) if (issuer) mbar_arrive_expect_tx(&full_bar[slot], 2 * bytes_per_cta);
        return leader_addr(&full_bar[slot]);
      };
      for (int kb = 0; kb < 4; ++kb) {                      // S: this tile's K' block [64 (key, head) x 64 k]
        const uint32_t fb = begin_unit(8192);
        if (issuer) tma_load_3d_2sm(ring + slot * VB_SLOT, &tmKp, fb, kb * 64, 0, cell_row0);
        next();
      }
      for (int kb = 0; kb < 4; ++kb) {                      // fc_o: Wo rows [rank*128, +128), k block kb
        const uint32_t fb = begin_unit(VB_SLOT);
        if (issuer) tma_load_2d_2sm(ring + slot * VB_SLOT, &tmWo, fb, kb * 64, static_cast<int>(rank) * 128);
        next();
      }
      for (int step = 0; step < 8; ++step) {                // FFN in MMA issue order: fc1(0), {fc1(c+1), fc2(c)}, fc2(3)
        const int is_fc2 = (step == 7) ? 1 : (step == 0 ? 0 : ((step & 1) ? 0 : 1));
        const int c = (step == 0) ? 0 : (step == 7 ? 3 : ((step & 1) ? (step + 1) >> 1 : (step >> 1) - 1));
        for (int kb = 0; kb < 4; ++kb) {
          const uint32_t fb = begin_unit(VB_SLOT);
          if (issuer) {
            if (!is_fc2) tma_load_2d_2sm(ring + slot * VB_SLOT, &tmW1, fb, kb * 64, c * 256 + static_cast<int>(rank) * 128);
            else tma_load_2d_2sm(ring + slot * VB_SLOT, &tmW2, fb, c * 256 + kb * 64, static_cast<int>(rank) * 128);
          }
          next();
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader) ===============================
    {   // the whole warp walks this block convergently (operands in uniform registers); lane 0 issues
      const bool issuer = (lane == 0);
      mbar_wait(in_bar, 0);                                  // this CTA's Q0 / V have landed
      if (issuer) mbar_arrive_cluster(leader_addr(in_both));
      if (rank == 0) {
        int slot = 0;
        uint32_t phase = 0;
        auto next = [&]() { if (++slot == VP_NS) { slot = 0; phase ^= 1; } };
        auto unit = [&](const uint8_t* a_sub, uint32_t d_col, uint32_t idesc, bool first) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(a_sub));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(ring + slot * VB_SLOT));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (issuer) umma_f16kind_2sm(tmem_base + d_col, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                             static_cast<uint32_t>(!(first && k == 0)));
          if (issuer) umma_commit_2sm(&empty_bar[slot]);
          next();
        };
        constexpr uint32_t idesc_s = umma_idesc_h16(256, 128);
        constexpr uint32_t idesc_w = umma_idesc_h16(256, 256);
        if (issuer) VB_STAMP(0);
        mbar_wait(in_both, 0);
        tc_fence_after();
        if (issuer) VB_STAMP(1);
        for (int kb = 0; kb < 4; ++kb) unit(bufA + kb * VB_SUB, 0, idesc_s, kb == 0);          // S -> columns [0, 128)
        if (issuer) umma_commit_2sm(s_done);
        if (issuer) VB_STAMP(2);
        mbar_wait(p_ready, 0);
        tc_fence_after();
        if (issuer) VB_STAMP(3);
        {
          constexpr uint32_t idesc_mn = umma_idesc_h16(256, 128) | (1u << 16);
          const uint64_t pdesc = umma_desc_sw128(smem_u32(sP));
          for (int h = 0; h < 4; ++h)                                                            // O_h -> columns [h*128, +128)
            if (issuer) umma_f16kind_2sm(tmem_base + h * 128, pdesc + static_cast<uint64_t>(h * 2), umma_desc_sw128(smem_u32(sV + h * 2048)),
                             idesc_mn, 0u);
          if (issuer) umma_commit_2sm(o_done);
        }
        mbar_wait(ctx_ready, 0);
        tc_fence_after();
        if (issuer) VB_STAMP(4);
        for (int kb = 0; kb < 4; ++kb) unit(bufB + kb * VB_SUB, TM_H, idesc_w, kb == 0);        // fc_o -> columns [256, 512)
        if (issuer) umma_commit_2sm(fco_done);
        if (issuer) VB_STAMP(5);
        mbar_wait(x_ready, 0);
        tc_fence_after();
        if (issuer) VB_STAMP(6);
        for (int step = 0; step < 8; ++step) {
          if (issuer) VB_STAMP(8 + step);
          const int is_fc2 = (step == 7) ? 1 : (step == 0 ? 0 : ((step & 1) ? 0 : 1));
          const int c = (step == 0) ? 0 : (step == 7 ? 3 : ((step & 1) ? (step + 1) >> 1 : (step >> 1) - 1));
          if (!is_fc2) {
            if (c >= 1) {      // both CTAs' epilogues have drained the (single) fc1 accumulator of chunk c - 1
              mbar_wait(hempty, static_cast<uint32_t>((c - 1) & 1));
              tc_fence_after();
            }
            for (int kb = 0; kb < 4; ++kb) unit(bufB + kb * VB_SUB, TM_H, idesc_w, kb == 0);
            if (issuer) umma_commit_2sm(hfull);
          } else {
            mbar_wait(sfull, static_cast<uint32_t>(c & 1));
            tc_fence_after();
            for (int kb = 0; kb < 4; ++kb) unit(bufA + kb * VB_SUB, TM_Y, idesc_w, c == 0 && kb == 0);
            if (issuer) umma_commit_2sm(sempty);
          }
        }
        if (issuer) umma_commit_2sm(y_done);
        if (issuer) VB_STAMP(7);
      }
    }
  } else {
    // =============================== epilogue (warps 2..9 of both CTAs) ===============================
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int sw = row & 7;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int et = threadIdx.x - 64;
    const uint32_t l_p_ready = leader_addr(p_ready), l_ctx_ready = leader_addr(ctx_ready), l_x_ready = leader_addr(x_ready),
                   l_hempty = leader_addr(hempty), l_sfull = leader_addr(sfull);
    if (et < 64) s_c[et] = from_h16(p.kvx[static_cast<long long>(cell_row0 + (et >> 2)) * p.kvx_pitch + 1024 + (et & 3)]);
    {
      const float* srcs[6] = {p.bo, p.ln1g, p.ln1b, p.b2, p.ln2g, p.ln2b};
#pragma unroll
      for (int i = 0; i < 6; ++i) s_par[i * 256 + et] = __ldg(srcs[i] + et);
#pragma unroll
      for (int i = 0; i < 4; ++i) s_par[6 * 256 + i * 256 + et] = __ldg(p.b1 + i * 256 + et);
    }
    named_bar(5, VB_EPI_WARPS * 32);

    // ---- softmax: this tile's scores are columns [rank*64, +64) of S (column = key*4 + head within the tile)
    float inv_sum[2];
    {
      if (et == 0) VB_STAMP(32);
      mbar_wait(s_done, 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(33);
      float s[64];
      {
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(trow + rank * 64, v0);
        tmem_ld_32x32(trow + rank * 64 + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) { s[j] = __uint_as_float(v0[j]); s[32 + j] = __uint_as_float(v1[j]); }
      }
      const float scale = 0.125f * 1.4426950408889634f;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        float x[16];
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          x[k] = (half == 0 ? s[k * 4 + hh] : s[k * 4 + 2 + hh]) + s_c[k * 4 + h];
          mx = fmaxf(mx, x[k]);
        }
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          x[k] = exp2f((x[k] - mx) * scale);
          sum += x[k];
        }
        inv_sum[hh] = 1.0f / sum;
        *reinterpret_cast<uint4*>(sP + row * 128 + (((2 * h) ^ sw) << 4)) = pack_chunk(&x[0]);
        *reinterpret_cast<uint4*>(sP + row * 128 + (((2 * h + 1) ^ sw) << 4)) = pack_chunk(&x[8]);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_p_ready);
    }

    // ---- ctx = O / rowsum: head h of this tile = TMEM columns [h*128 + rank*64, +64)
    {
      if (et == 0) VB_STAMP(34);
      mbar_wait(o_done, 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(35);
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int h = 2 * half + hh;
        uint32_t v0[32], v1[32];
        tmem_ld_32x32(trow + h * 128 + rank * 64, v0);
        tmem_ld_32x32(trow + h * 128 + rank * 64 + 32, v1);
        tmem_ld_wait();
        uint8_t* dst = bufB + h * VB_SUB + row * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float g[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = __uint_as_float(j < 4 ? v0[j * 8 + e] : v1[(j - 4) * 8 + e]) * inv_sum[hh];
          *reinterpret_cast<uint4*>(dst + ((j ^ sw) << 4)) = pack_chunk(g);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_ctx_ready);
    }

    // ---- X = LN1(Q0 + A + bo)
    {
      if (et == 0) VB_STAMP(36);
      mbar_wait(in_bar, 0);
      mbar_wait(fco_done, 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(37);
      const int n0 = half * 128;
      ln_epilogue(trow + TM_H + n0, bufA + 2 * half * VB_SUB, bufB + 2 * half * VB_SUB, s_par + n0, s_par + 256 + n0,
                  s_par + 512 + n0, s_ln, half, row, quad, p.eps, [](int, const uint4&) {},
                  (p.times != nullptr && et == 0) ? p.times + static_cast<long long>(blockIdx.x) * 64 + 61 : nullptr);
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_x_ready);
    }

    // ---- H_c = relu(fc1 chunk + b1), 4 chunks of 256 hidden units; this thread owns 128 of them = sub-tiles 2*half, 2*half+1
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      if (et == 0) VB_STAMP(40 + 2 * c);
      mbar_wait(hfull, static_cast<uint32_t>(c & 1));
      tc_fence_after();
      if (et == 0) VB_STAMP(41 + 2 * c);
      float f[128];
      tmem_ld_128(trow + TM_H + half * 128, f);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_hempty);          // drained: fc1(c + 1) may overwrite the accumulator
      if (c >= 1) mbar_wait(sempty, static_cast<uint32_t>((c - 1) & 1));   // fc2(c - 1) has read the H buffer
      const float* bias = s_par + 6 * 256 + c * 256 + half * 128;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 b0 = *reinterpret_cast<const float4*>(bias + j * 8);
        const float4 b1 = *reinterpret_cast<const float4*>(bias + j * 8 + 4);
        float g[8];
        g[0] = fmaxf(f[j * 8] + b0.x, 0.0f); g[1] = fmaxf(f[j * 8 + 1] + b0.y, 0.0f);
        g[2] = fmaxf(f[j * 8 + 2] + b0.z, 0.0f); g[3] = fmaxf(f[j * 8 + 3] + b0.w, 0.0f);
        g[4] = fmaxf(f[j * 8 + 4] + b1.x, 0.0f); g[5] = fmaxf(f[j * 8 + 5] + b1.y, 0.0f);
        g[6] = fmaxf(f[j * 8 + 6] + b1.z, 0.0f); g[7] = fmaxf(f[j * 8 + 7] + b1.w, 0.0f);
        *reinterpret_cast<uint4*>(bufA + (2 * half + (j >> 3)) * VB_SUB + row * 128 + (((j & 7) ^ sw) << 4)) = pack_chunk(g);
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(l_sfull);
    }

    // ---- Y = LN2(X + Y + b2) in place of X, then the token mean
    {
      if (et == 0) VB_STAMP(38);
      mbar_wait(y_done, 0);
      tc_fence_after();
      if (et == 0) VB_STAMP(39);
      const int n0 = half * 128;
      h16* ytok = (p.y_tokens != nullptr && row < p.L) ? p.y_tokens + ((static_cast<long long>(mod) * p.B + env) * p.L + row) * 256 + n0 : nullptr;
      ln_epilogue(trow + TM_Y + n0, bufB + 2 * half * VB_SUB, bufB + 2 * half * VB_SUB, s_par + 768 + n0, s_par + 1024 + n0,
                  s_par + 1280 + n0, s_ln, half, row, quad, p.eps, [ytok](int col, const uint4& yq) {
                    if (ytok != nullptr) *reinterpret_cast<uint4*>(ytok + col) = yq;
                  }, (p.times != nullptr && et == 0) ? p.times + static_cast<long long>(blockIdx.x) * 64 + 59 : nullptr);
      if (et == 0) VB_STAMP(57);
      named_bar(5, VB_EPI_WARPS * 32);
      if (et == 0) VB_STAMP(58);
      const uint8_t* colp = bufB + (et >> 6) * VB_SUB + (et & 7) * 2;
      const int chunk = (et & 63) >> 3;
      float acc4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      int r = 0;
      for (; r + 8 <= p.L; r += 8) {
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = from_h16(*reinterpret_cast<const h16*>(colp + (r + k) * 128 + ((chunk ^ k) << 4)));
#pragma unroll
        for (int k = 0; k < 8; ++k) acc4[k & 3] += x[k];
      }
      // tail: r is a multiple of 8 here, so row r + k has k = (r + k) & 7 and goes to accumulator k & 3 -- static indices
      // (indexing acc4 with the run-time row number had put the four accumulators in local memory)
#pragma unroll
      for (int k = 0; k < 7; ++k)
        if (r + k < p.L) acc4[k & 3] += from_h16(*reinterpret_cast<const h16*>(colp + (r + k) * 128 + ((chunk ^ k) << 4)));
      const float acc = (acc4[0] + acc4[1]) + (acc4[2] + acc4[3]);
      p.out[static_cast<long long>(env) * p.out_pitch + mod * 256 + et] = to_h16(acc / static_cast<float>(p.L));
      if (et == 0) VB_STAMP(56);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be reading this CTA's half of a weight block / using its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

}  // namespace

void vla_block_make_plan(const VlaBlock& d, VlaBlockPlan* plan) {
  RVB_CHECK(d.L >= 1 && d.L <= 128, "vla_block: 1 <= L <= 128 query tokens per environment");
  RVB_CHECK(d.B >= 1 && d.q0 != nullptr && d.kvx != nullptr && d.wo != nullptr && d.w1 != nullptr && d.w2 != nullptr && d.out != nullptr,
            "vla_block: null operand");
  RVB_CHECK(d.kvx_pitch == 1288, "vla_block: kvx rows are [K'(4 x 256) | c(8) | V(256)]");
  plan->d = d;
  const uint64_t q_rows = static_cast<uint64_t>(d.q_shared ? 1 : d.B) * d.L;
  tma_encode_2d_h16(&plan->tmQ0, d.q0, 256, q_rows, 512, 64, 128);
  {
    const uint64_t dims[3] = {256, 4, static_cast<uint64_t>(2) * d.B * 16};
    const uint64_t strides[2] = {512, static_cast<uint64_t>(d.kvx_pitch) * 2};
    const uint32_t box[3] = {64, 4, 16};
    tma_encode_nd_h16(&plan->tmKp, d.kvx, 3, dims, strides, box);
  }
  tma_encode_2d_h16(&plan->tmV, d.kvx, static_cast<uint64_t>(d.kvx_pitch), static_cast<uint64_t>(2) * d.B * 16,
                    static_cast<uint64_t>(d.kvx_pitch) * 2, 64, 16);
  tma_encode_2d_h16(&plan->tmWo, d.wo, 256, 256, 512, 64, 128);
  tma_encode_2d_h16(&plan->tmW1, d.w1, 256, 1024, 512, 64, 128);
  tma_encode_2d_h16(&plan->tmW2, d.w2, 1024, 256, 2048, 64, 128);
  plan->valid = true;
}

// 0: environment default (ROBOVLN_VLA_PAIR, default 1 = CTA-pair kernel); 1: one CTA per tile; 2: CTA pair per environment
int g_vla_variant = 0;

static bool use_pair_kernel() {
  if (g_vla_variant != 0) return g_vla_variant == 2;
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_VLA_PAIR");
    v = (e != nullptr && std::strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}

void vla_block_launch(const VlaBlockPlan& plan, cudaStream_t s) {
  RVB_CHECK(plan.valid, "vla_block: plan not built");
  const bool pair = use_pair_kernel();
  static PerDeviceOnce attr_once[2];
  if (attr_once[pair ? 1 : 0].first()) {
    if (pair) RVB_CUDA(cudaFuncSetAttribute(vla_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_SMEM));
    else RVB_CUDA(cudaFuncSetAttribute(vla_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, VB_SMEM));
  }
  const VlaBlock& d = plan.d;
  VlaBlockParams p;
  std::memset(&p, 0, sizeof(p));
  p.B = d.B; p.L = d.L; p.q_shared = d.q_shared;
  p.kvx = d.kvx; p.kvx_pitch = d.kvx_pitch;
  p.bo = d.bo; p.b1 = d.b1; p.b2 = d.b2; p.ln1g = d.ln1g; p.ln1b = d.ln1b; p.ln2g = d.ln2g; p.ln2b = d.ln2b;
  p.eps = d.eps; p.out = d.out; p.out_pitch = d.out_pitch; p.y_tokens = d.y_tokens;
  static const char* renv = std::getenv("ROBOVLN_VLA_ROTATE");
  p.rotate = (renv != nullptr && std::strcmp(renv, "1") == 0) ? 1 : 0;   // measured: no effect on B200 (the L2 serves same-line readers fine)
  // ROBOVLN_VLA_TIMES=<file>: per-CTA phase time stamps (SM clocks) of every launch are written to <file> (diagnostics;
  // synchronises the stream)
  static const char* tenv = std::getenv("ROBOVLN_VLA_TIMES");
  long long* tbuf = nullptr;
  const size_t tbytes = static_cast<size_t>(d.B) * 2 * 64 * sizeof(long long);
  if (tenv != nullptr) {
    RVB_CUDA(cudaMalloc(&tbuf, tbytes));
    RVB_CUDA(cudaMemsetAsync(tbuf, 0, tbytes, s));
  }
  p.times = tbuf;
  if (pair) {
    cudaLaunchConfig_t cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * d.B, 1, 1);
    cfg.blockDim = dim3(VB_THREADS, 1, 1);
    cfg.dynamicSmemBytes = VB_SMEM;
    cfg.stream = s;
    cudaLaunchAttribute attr[3];
    int na = 0;
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
    if (g_launch_prio != 0) {
      attr[na].id = cudaLaunchAttributePriority;
      attr[na].val.priority = g_launch_prio;
      ++na;
    }
    if (use_pdl()) {
      attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    RVB_CUDA(cudaLaunchKernelEx(&cfg, vla_pair_kernel, plan.tmQ0, plan.tmKp, plan.tmV, plan.tmWo, plan.tmW1, plan.tmW2, p));
  } else {
    launch_k(vla_block_kernel, dim3(d.B, 2), dim3(VB_THREADS), VB_SMEM, s, plan.tmQ0, plan.tmKp, plan.tmV, plan.tmWo, plan.tmW1,
             plan.tmW2, p);
  }
  RVB_CUDA(cudaGetLastError());
  if (tbuf != nullptr) {
    std::vector<long long> host(static_cast<size_t>(d.B) * 2 * 64);
    RVB_CUDA(cudaStreamSynchronize(s));
    RVB_CUDA(cudaMemcpy(host.data(), tbuf, tbytes, cudaMemcpyDeviceToHost));
    RVB_CUDA(cudaFree(tbuf));
    FILE* f = std::fopen(tenv, "a");
    if (f != nullptr) {
      std::fprintf(f, "# launch B=%d L=%d pair=%d\n", d.B, d.L, pair ? 1 : 0);
      for (int c = 0; c < d.B * 2; ++c) {
        std::fprintf(f, "%d", c);
        for (int i = 0; i < 64; ++i) std::fprintf(f, ",%lld", host[static_cast<size_t>(c) * 64 + i]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
  }
}

}  // namespace rvb
