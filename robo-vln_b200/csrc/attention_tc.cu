// BERT self-attention on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), one CTA per
// (instruction row, head), L <= 128 keys.
//
// Reference: transformers BertSelfAttention as called by seq2seq_highlevel_cma.py:192-195 (no
// attention mask): ctx = softmax(Q K^T / 8) V per head.
//
//   1. TMA loads the head's Q [128 x 64], K [LP x 64] and V [LP x 64] slices straight out of the packed
//      QKV activation matrix [R*L, 2304] into 128B-swizzled shared memory (LP = L rounded up to 16; rows
//      past the sequence are the next sequence's or zero fill and are masked / discarded).
//   2. S = Q K^T: tcgen05.mma M=128, N=LP, K=64 (both operands K-major) -> TMEM columns [0, LP).
//   3. Softmax straight from TMEM, one thread per query row (tcgen05.ld), two passes (max, then
//      exp / sum); the un-normalised probabilities are written as 16-bit values into a swizzled
//      K-major shared-memory tile = the A operand of the second product.
//   4. O = P V: tcgen05.mma M=128, N=64, K=LP with V consumed as an MN-MAJOR B operand -- the [key, dim]
//      tile TMA delivered is exactly the canonical MN-major SW128 layout, so V is never transposed.
//   5. O rows are read back from TMEM, scaled by 1/sum and stored (128 B per row).
//
// Selected with ROBOVLN_ATTN=tc (and exercised by the kernel tests through rvb_bert_attention_tc).  The default
// stays the warp-level mma.sync kernel of attention.cu: with one (row, head) per CTA this kernel is a serial
// TMA -> MMA -> softmax -> MMA -> store latency chain and measures 4 % slower on the BERT stream at L = 80
// (1.75 vs 1.68 ms); batching several heads per CTA behind a pipeline is the obvious next step.
#include "common.cuh"
#include "rvb.h"

#include <cstdlib>
#include <cstring>

namespace rvb {

namespace {

constexpr int AT_THREADS = 128;
constexpr int AT_HD = 64;
// smem: Q 16 KB | K 16 KB | V 16 KB | P 2 x 16 KB | barriers
constexpr int AT_SMEM = 5 * 16384 + 1024 /*align slack*/ + 64;

RVB_DEVICE void tmem_ld_32x32_x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

__global__ void __launch_bounds__(AT_THREADS) bert_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                   const __grid_constant__ CUtensorMap tmKV,
                                                                   h16* __restrict__ ctx, int L, int LP, int heads) {
  extern __shared__ uint8_t at_raw[];
  uint8_t* smem = at_raw + ((1024u - (smem_u32(at_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS / STS)
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;
  uint8_t* sV = smem + 2 * 16384;
  uint8_t* sP = smem + 3 * 16384;   // two K blocks of 128 rows x 128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 5 * 16384);   // [0] loads, [1] S ready, [2] O ready
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  const int head = blockIdx.x, row = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int H = heads * AT_HD;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmKV);
    for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  RVB_PDL_PROLOGUE();

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384u + 2u * static_cast<uint32_t>(LP) * 128u);
    tma_load_2d(sQ, &tmQ, &bars[0], head * AT_HD, row * L);
    tma_load_2d(sK, &tmKV, &bars[0], H + head * AT_HD, row * L);
    tma_load_2d(sV, &tmKV, &bars[0], 2 * H + head * AT_HD, row * L);
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    // S[128, LP] = Q K^T   (K = 64: four k-steps of 16, 32 B apart inside the 128 B swizzle row)
    const uint32_t idesc = umma_idesc_h16(128, static_cast<uint32_t>(LP));
    const uint64_t adesc = umma_desc_sw128(smem_u32(sQ)), bdesc = umma_desc_sw128(smem_u32(sK));
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_f16kind(tmem_base, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                   static_cast<uint32_t>(k != 0));
    umma_commit(&bars[1]);
  }

  // ---- softmax, one thread per query row (TMEM lane = row) ----
  const int t = threadIdx.x;
  const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  float mx = -INFINITY;
  for (int c0 = 0; c0 < LP; c0 += 16) {
    uint32_t v[16];
    __syncwarp();
    tmem_ld_32x32_x16(trow + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (c0 + j < L) mx = fmaxf(mx, __uint_as_float(v[j]));
  }
  const float scale = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)
  float sum = 0.0f;
  const int sw = t & 7;
  for (int c0 = 0; c0 < LP; c0 += 16) {
    uint32_t v[16];
    __syncwarp();
    tmem_ld_32x32_x16(trow + c0, v);
    tmem_ld_wait();
    float p[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      p[j] = (c0 + j < L) ? exp2f((__uint_as_float(v[j]) - mx) * scale) : 0.0f;
      sum += p[j];
    }
    // columns c0 .. c0+15 of row t = two 16-byte chunks of K block c0 / 64
    uint8_t* prow = sP + (c0 >> 6) * 16384 + t * 128;
    const int ch = (c0 & 63) >> 3;
    uint4 q0, q1;
    q0.x = pack_h2(p[0], p[1]); q0.y = pack_h2(p[2], p[3]); q0.z = pack_h2(p[4], p[5]); q0.w = pack_h2(p[6], p[7]);
    q1.x = pack_h2(p[8], p[9]); q1.y = pack_h2(p[10], p[11]); q1.z = pack_h2(p[12], p[13]); q1.w = pack_h2(p[14], p[15]);
    *reinterpret_cast<uint4*>(prow + ((ch ^ sw) << 4)) = q0;
    *reinterpret_cast<uint4*>(prow + (((ch + 1) ^ sw) << 4)) = q1;
  }
  fence_proxy_async();   // P (generic-proxy stores) -> UMMA (async proxy)
  tc_fence_before();
  __syncthreads();

  if (threadIdx.x == 0) {
    tc_fence_after();
    // O[128, 64] = P V   (K = LP; V is the MN-major B operand: [key, dim] rows of 128 B, 8-key groups 1024 B apart)
    const uint32_t idesc = umma_idesc_h16(128, 64) | (1u << 16);
    const uint64_t vdesc = umma_desc_sw128(smem_u32(sV));
    for (int ks = 0; ks < LP / 16; ++ks) {
      const uint64_t adesc = umma_desc_sw128(smem_u32(sP + (ks >> 2) * 16384)) + static_cast<uint64_t>((ks & 3) * 2);
      umma_f16kind(tmem_base + 128, adesc, vdesc + static_cast<uint64_t>(ks * 128), idesc, static_cast<uint32_t>(ks != 0));
    }
    umma_commit(&bars[2]);
  }
  mbar_wait(&bars[2], 0);
  tc_fence_after();
  {
    const float inv = 1.0f / sum;
    h16* out = ctx + (static_cast<long long>(row) * L + t) * H + head * AT_HD;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t v[32];
      __syncwarp();
      tmem_ld_32x32(trow + 128 + half * 32, v);
      tmem_ld_wait();
      if (t < L) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 q;
          q.x = pack_h2(__uint_as_float(v[8 * j]) * inv, __uint_as_float(v[8 * j + 1]) * inv);
          q.y = pack_h2(__uint_as_float(v[8 * j + 2]) * inv, __uint_as_float(v[8 * j + 3]) * inv);
          q.z = pack_h2(__uint_as_float(v[8 * j + 4]) * inv, __uint_as_float(v[8 * j + 5]) * inv);
          q.w = pack_h2(__uint_as_float(v[8 * j + 6]) * inv, __uint_as_float(v[8 * j + 7]) * inv);
          *reinterpret_cast<uint4*>(out + half * 32 + j * 8) = q;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

bool use_tc_attention() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_ATTN");
    v = (e != nullptr && std::strcmp(e, "tc") == 0) ? 1 : 0;   // measured: the mma.sync kernel is faster at L = 80
  }
  return v == 1;
}

// qkv [R*L, 3*heads*64] h16 (Q | K | V column blocks), ctx [R*L, heads*64] h16; 1 <= L <= 128.
void bert_self_attention_tc(const h16* qkv, h16* ctx, int R, int L, int heads, cudaStream_t s) {
  RVB_CHECK(L >= 1 && L <= 128, "tcgen05 attention: 1 <= L <= 128");
  RVB_CHECK((reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(ctx) & 15) == 0, "attention: alignment");
  const int LP = (L + 15) / 16 * 16;
  const uint64_t rows = static_cast<uint64_t>(R) * L, cols = 3ull * heads * AT_HD;
  CUtensorMap tmQ, tmKV;
  tma_encode_2d_h16(&tmQ, qkv, cols, rows, cols * 2, AT_HD, 128);
  tma_encode_2d_h16(&tmKV, qkv, cols, rows, cols * 2, AT_HD, static_cast<uint32_t>(LP));
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    RVB_CUDA(cudaFuncSetAttribute(bert_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AT_SMEM));
  }
  launch_k(bert_attn_tc_kernel, dim3(heads, R), dim3(AT_THREADS), AT_SMEM, s, tmQ, tmKV, ctx, L, LP, heads);
  RVB_CUDA(cudaGetLastError());
}

}  // namespace rvb
