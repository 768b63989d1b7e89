// tcgen05 / TMEM / TMA implicit-GEMM kernel for sm_100a.
//
// One persistent, warp-specialised kernel covers every dense contraction of the HCM policy
// forward (reference call sites: torchvision ResNet-50 convs behind
// robo_vln_baselines/models/encoders/resnet_encoders.py:189-237, the DDPPO GroupNorm ResNet
// convs of habitat_baselines/rl/ddppo/policy/resnet.py:65-77, every nn.Linear of BERT and of
// robo_vln_baselines/models/transformer/transformer.py, the kv/linear heads of
// seq2seq_highlevel_cma.py:83-115 and the LSTM input projections):
//
//   out[m, n] = act( sum_{tap, c} A(m, tap, c) * W[n, tap*Cin + c] + bias[n] + res[m, n] )
//
// * A operand: NHWC 16-bit activations read straight from the tensor by TMA -- no im2col
//   buffer.  For a KHxKW filter the K loop walks (tap, 64-channel block); each step is one
//   4-D box load (64 ch, Wo, th, nb) whose start coordinate is shifted by the tap offset,
//   with TMA out-of-bounds zero fill providing the padding and elementStrides providing the
//   convolution stride.  Because an M tile is a set of full output rows, the box lands in
//   shared memory exactly as the 128-row, K-major, 128B-swizzled tile UMMA expects.
//   "Window" mode (the 7x7 stride-2 RGB stem): the input is a zero-padded NHW8 image and the
//   tensor map strides the W dimension by 2 pixels (32 B) while each row of the box reads 64
//   contiguous elements = 8 pixels x 8 channels, i.e. overlapping windows: one tap ROW of the
//   7x7 filter per K block, 7 K blocks in all.
// * B operand: weights [Cout, KH*KW*Cin], K-major, 2-D TMA boxes (64, BN).
// * MMA: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16, fp32 accumulators in TMEM,
//   double-buffered (2 x BN columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
// * Roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
//   warps 2..9 = epilogue in two groups of four (one warp per TMEM lane quadrant); the groups
//   take alternate 128-byte column chunks: tcgen05.ld -> bias / residual / ReLU|GELU ->
//   128B-swizzled smem staging -> TMA store (coalesced, clipped at the tensor edges).
// * Epilogue variants (third template parameter): 0 = bias / residual / activation; 1 = LayerNorm folded into the
//   store (two passes over the TMEM accumulator with tcgen05.ld / tcgen05.st, row statistics exchanged between the
//   N/256 CTAs of a row block over DSMEM or global memory); 2 = GroupNorm folded into the store for feature maps of
//   <= 64 pixels (whole samples per tile: segmented warp-shuffle statistics) -- the depth trunk's layers 3-4.
// * The RGB stem runs here too: a row-pair-interleaved padded image and an overlapping-window tensor map put two
//   filter rows in every 64-wide K block ("window == 2").
#include "common.cuh"
#include "rvb.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

namespace rvb {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int NUM_EPI_WARPS = 8;
constexpr int NUM_THREADS = 64 + NUM_EPI_WARPS * 32;  // 320
constexpr int STAGING_BYTES = 2 * 16384;              // one 128 x 128B chunk buffer per epilogue group
constexpr int SMEM_STAGE_BUDGET = 196608;             // 192 KiB of pipeline stages

// CTAS = 1: one CTA owns a 128 x BN tile.  CTAS = 2: a CTA pair (2-CTA cluster) owns a 256 x BN
// tile; each CTA stages its 128 rows of A and HALF of B (BN/2 rows) per K block, which doubles
// the FLOPs per byte pulled into each SM -- the per-SM L2 ingress is what bounds the 1-CTA form.
template <int BN, int CTAS = 1>
struct Cfg {
  static constexpr int B_STAGE_BYTES = (BN / CTAS) * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_STAGE_BYTES + B_STAGE_BYTES;
  static constexpr int STAGES = (SMEM_STAGE_BUDGET / STAGE_BYTES) > 8 ? 8 : (SMEM_STAGE_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BN;  // 128, 256 or 512: powers of two >= 32
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

// phase stamps for tools/gemm_timeline.py: compiled in only with -DRVB_GEMM_STAMPS=1 (ROBOVLN_BUILD_STAMPS=1 python build.py).
// Only stamps that follow an mbarrier wait are meaningful: clock reads float freely among independent ALU instructions.
#if defined(RVB_GEMM_STAMPS) && RVB_GEMM_STAMPS
#define GT_STAMP(i)                                                                                  \
  do {                                                                                               \
    if (p.times != nullptr) p.times[static_cast<long long>(blockIdx.x) * 16 + (i)] = clock64();      \
  } while (0)
#else
#define GT_STAMP(i) do { } while (0)
#endif

RVB_DEVICE void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// bias + residual + activation on 32 consecutive columns of one output row
RVB_DEVICE void epilogue_math(float (&f)[32], const GemmTcParams& p, int n, const h16* res_row) {
  if (p.bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      if (n + j < p.N) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n + j));
        f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
      }
    }
  }
  if (res_row != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (n + j < p.N) {
        const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(res_row + n + j));
        float2 t;
        t = unpack_h2(r4.x); f[j] += t.x; f[j + 1] += t.y;
        t = unpack_h2(r4.y); f[j + 2] += t.x; f[j + 3] += t.y;
        t = unpack_h2(r4.z); f[j + 4] += t.x; f[j + 5] += t.y;
        t = unpack_h2(r4.w); f[j + 6] += t.x; f[j + 7] += t.y;
      }
    }
  }
  if (p.act == ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
  } else if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
  }
}

// epilogue_math with the bias of the 32 columns arriving lane-distributed (lane l holds bias[n + l]).  With > 200 KB of
// shared memory carved out the L1 holds nothing, so the per-thread bias loads of epilogue_math are L2 round trips on
// the epilogue's critical path: measured (tools/gemm_timeline.py + ROBOVLN_EPI_DEBUG=2) 3.8k of the 7.7k epilogue cycles
// of a 128 x 256 bias-only tile.  The lane-distributed copy is fetched once per tile BEFORE the accumulator is waited for.
RVB_DEVICE void epilogue_math_b(float (&f)[32], const GemmTcParams& p, int n, const h16* res_row, float bias_lane,
                                bool relu_in_pack = false) {   // relu_in_pack: the caller's 16-bit pack applies the ReLU (F2FP.RELU)
#pragma unroll
  for (int j = 0; j < 32; ++j) f[j] += __shfl_sync(0xffffffffu, bias_lane, j);
  if (res_row != nullptr) {
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
      if (n + j < p.N) {
        const uint4 r4 = __ldg(reinterpret_cast<const uint4*>(res_row + n + j));
        float2 t;
        t = unpack_h2(r4.x); f[j] += t.x; f[j + 1] += t.y;
        t = unpack_h2(r4.y); f[j + 2] += t.x; f[j + 3] += t.y;
        t = unpack_h2(r4.z); f[j + 4] += t.x; f[j + 5] += t.y;
        t = unpack_h2(r4.w); f[j + 6] += t.x; f[j + 7] += t.y;
      }
    }
  }
  if (p.act == ACT_RELU) {
    if (!relu_in_pack) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.0f);
    }
  } else if (p.act == ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = gelu_fast(f[j]);
  }
}

template <int BN, int CTAS, int EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR,
               const GemmTcParams p) {
  constexpr bool LN = (EPI == 1);   // LayerNorm folded into the store
  constexpr bool GN = (EPI == 2);   // GroupNorm folded into the store
  using C = Cfg<BN, CTAS>;
  constexpr int STAGES = C::STAGES;       // ring capacity; p.nstages (<= STAGES) are in use
  const int nstages = p.nstages;
  const uint32_t cta_rank = (CTAS == 2) ? cluster_ctarank() : 0u;   // 0 = leader of the pair
  // LayerNorm mode (CTAS == 1): the n_tiles CTAs of one 128-row block form a cluster; CTA `ln_rank` owns
  // column tile ln_rank and all CTAs of a cluster walk the same sequence of row blocks
  const int ncl = (LN && !p.ln_xchg) ? p.n_tiles : 1;   // cluster size (DSMEM exchange); 1 for the global-memory exchange
  const uint32_t ln_rank = (ncl > 1) ? cluster_ctarank() : 0u;
  const int unit = (CTAS == 2) ? static_cast<int>(blockIdx.x >> 1)
                   : (ncl > 1 ? static_cast<int>(blockIdx.x / ncl) * p.n_tiles + static_cast<int>(ln_rank)
                              : static_cast<int>(blockIdx.x));
  const int num_units = (CTAS == 2) ? static_cast<int>(gridDim.x >> 1)
                        : (ncl > 1 ? static_cast<int>(gridDim.x / ncl) * p.n_tiles : static_cast<int>(gridDim.x));

  extern __shared__ uint8_t smem_raw[];
  if (threadIdx.x == 0) GT_STAMP(0);
  __shared__ float gn_x[GN ? 2 * 8 * 16 : 1];   // GroupNorm: (sum, sumsq) x 8 groups per epilogue warp, double-buffered
  // 1024-byte alignment by pointer arithmetic ON the __shared__ array: rounding the address through an integer made every
  // later access a generic LD.E / ST.E instead of LDS / STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  // resident-B mode: [nstages A buffers][num_kb B blocks]; otherwise [STAGES A buffers][STAGES B buffers]
  uint8_t* smem_b = smem + (p.b_res ? nstages : STAGES) * A_STAGE_BYTES;
  uint8_t* smem_c = smem + STAGES * C::STAGE_BYTES;  // 2 x 16 KiB staging, 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + STAGING_BYTES);
  uint64_t* full_bar = bars;                     // [STAGES]
  uint64_t* empty_bar = bars + STAGES;           // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;       // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* rbar_base = bars + 2 * STAGES + 5;   // [8] residual-slice barriers, one per epilogue warp
  uint64_t* ln_bar = rbar_base + NUM_EPI_WARPS;  // [2] LayerNorm exchange: row sums / centred sums of squares
  // residual-prefetch mode: the pipeline runs with fewer stages and the B buffers of the unused
  // stages hold eight 4 KiB residual slices (32 rows x 128 B, one per epilogue warp)
  uint8_t* res_slices = smem_b + nstages * C::B_STAGE_BYTES;

  // the warp index through a shuffle: provably warp-uniform for the compiler, so that the role branches below are
  // uniform branches and the single-lane TMA / MMA issue loops can keep their operands in uniform registers
  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    if (p.res_tma) tma_prefetch_desc(&tmR);
    for (int i = 0; i < NUM_EPI_WARPS; ++i) mbar_init(&rbar_base[i], 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], NUM_EPI_WARPS * CTAS);  // one arrive per epilogue warp (of both CTAs)
      if (LN) mbar_init(&ln_bar[i], static_cast<uint32_t>(ncl) * NUM_EPI_WARPS * 32);   // every epilogue thread of the cluster ([0] in use)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CTAS == 2) tmem_alloc_2sm<C::TMEM_COLS>(tmem_slot);
    else tmem_alloc<C::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2 || ncl > 1) cluster_sync_all();   // peer barriers are initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above overlapped the previous kernel's tail; its outputs (our A operand and
  // residual) and our output buffer (which it may still be reading) are safe only from here on
  pdl_launch_dependents();
  pdl_wait();
  if (threadIdx.x == 0) GT_STAMP(1);

  // work units: 128 x BN tiles (CTAS == 1) or 256 x BN pair tiles (CTAS == 2, this CTA = rows rank*128..)
  const int m_units = (p.m_tiles + CTAS - 1) / CTAS;
  const int total_tiles = m_units * p.n_tiles;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    // The whole warp walks the loop (convergent code: coordinates live in uniform registers); lane 0 issues.  With the
    // loop inside `if (lane == 0)` every TMA instruction was wrapped in an elect / R2UR.BROADCAST waterfall and a k block
    // took ~480 cycles to issue.
    {
      const bool issuer = (lane == 0);
      int stage = 0;
      uint32_t phase = 0;
      if (CTAS == 1 && p.b_res && unit < total_tiles) {   // the whole weight matrix of this (single) N tile, once
        if (issuer) mbar_arrive_expect_tx(&rbar_base[0], static_cast<uint32_t>(p.num_kb) * C::B_STAGE_BYTES);
        int tapc = 0, cbr = 0;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          if (issuer) tma_load_2d(smem_b + kb * C::B_STAGE_BYTES, &tmB, &rbar_base[0], tapc + cbr * BLOCK_K, 0);
          if (++cbr == p.cblocks) {
            cbr = 0;
            tapc += p.Cin;
          }
        }
      }
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        const int mu = (p.n_tiles == 1) ? tile : tile / p.n_tiles;   // no division on the single-N-tile launches (one k block per tile there)
        const int nt = tile - mu * p.n_tiles;
        const int mt = mu * CTAS + static_cast<int>(cta_rank);   // may be one past the end for the peer: OOB -> zeros
        int img = 0, h0 = 0;
        if (!p.plain) {
          if (p.nb == 1) {
            img = mt / p.tiles_per_img;
            h0 = (mt - img * p.tiles_per_img) * p.th;
          } else {
            img = mt * p.nb;
          }
        }
        // The K walk (filter tap (r, s) x 64-channel block) is kept in running counters: this loop is ONE thread, every
        // instruction of it is issued at dependent-chain latency, and the two integer divisions per k block it used to
        // do (kb / cblocks, tap / KW) made a k block cost ~640 cycles to issue -- more than its 128 MMA cycles at
        // BN = 64, so the narrow convs of the trunk fronts were bound by their own producer.
        int tap = 0, cb = 0, r = 0, sx = 0, tapcol = 0;
        const int hbase = h0 * p.stride - p.pad;
        const int ncol = nt * BN + ((CTAS == 2) ? static_cast<int>(cta_rank) * (BN / 2) : 0);
        const int mrow = mt * BLOCK_M;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * C::B_STAGE_BYTES;
          const int ccol = cb * BLOCK_K;
          if (CTAS == 2) {
            // both CTAs' bytes are credited to the LEADER's full barrier, which the leader arms
            const uint32_t full_leader = mapa_u32(smem_u32(&full_bar[stage]), 0);
            if (issuer) {
              if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * (p.a_bytes + C::B_STAGE_BYTES));
              if (p.plain) tma_load_4d_2sm(sa, &tmA, full_leader, ccol, mrow, 0, 0);
              else tma_load_4d_2sm(sa, &tmA, full_leader, ccol, sx - p.pad, hbase + r, img);
              tma_load_2d_2sm(sb, &tmB, full_leader, tapcol + ccol, ncol);
            }
          } else if (issuer) {
            mbar_arrive_expect_tx(&full_bar[stage], p.a_bytes + (p.b_res ? 0u : static_cast<uint32_t>(C::B_STAGE_BYTES)));
            if (p.plain) tma_load_4d(sa, &tmA, &full_bar[stage], ccol, mrow, 0, 0);
            else if (p.window2) tma_load_4d(sa, &tmA, &full_bar[stage], 0, 0, h0 + tap, img);   // row pair h0 + tap = filter rows 2 tap, 2 tap + 1
            else tma_load_4d(sa, &tmA, &full_bar[stage], ccol, sx - p.pad, hbase + r, img);
            if (!p.b_res) tma_load_2d(sb, &tmB, &full_bar[stage], tapcol + ccol, ncol);
          }
          if (issuer && tile == unit && kb == 0) GT_STAMP(2);
          if (++cb == p.cblocks) {      // next filter tap
            cb = 0;
            ++tap;
            tapcol += p.Cin;
            if (++sx == p.KW) {
              sx = 0;
              ++r;
            }
          }
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (issuer && tile == unit) GT_STAMP(3);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (cta_rank == 0) {   // in a pair only the leader issues MMAs; the warp walks the loop convergently, lane 0 issues
      const bool issuer = (lane == 0);
      constexpr uint32_t idesc = umma_idesc_h16(BLOCK_M * CTAS, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      // descriptors of slot 0; a slot's descriptor is that plus its byte offset >> 4 (the address field holds 14 bits)
      const uint64_t adesc0 = umma_desc_sw128(smem_u32(smem_a));
      const uint64_t bdesc0 = umma_desc_sw128(smem_u32(smem_b));
      if (CTAS == 1 && p.b_res && unit < total_tiles) mbar_wait(&rbar_base[0], 0);   // resident weights have landed
      for (int tile = unit; tile < total_tiles; tile += num_units) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (issuer && tile == unit && kb == 0) GT_STAMP(4);
          const uint64_t adesc = adesc0 + static_cast<uint64_t>(stage * (A_STAGE_BYTES >> 4));
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>((p.b_res ? kb : stage) * (C::B_STAGE_BYTES >> 4));
          if (issuer) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advance 16 elements = 32 B along K inside the 128B swizzle atom: +2 in the >>4 address field
              if (CTAS == 2)
                umma_f16kind_2sm(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                                 static_cast<uint32_t>((kb | k) != 0));
              else
                umma_f16kind(d_tmem, adesc + static_cast<uint64_t>(k * 2), bdesc + static_cast<uint64_t>(k * 2), idesc,
                             static_cast<uint32_t>((kb | k) != 0));
            }
            // frees the smem slot (in both CTAs of a pair) once these MMAs retire
            if (CTAS == 2) umma_commit_2sm(&empty_bar[stage]);
            else umma_commit(&empty_bar[stage]);
          }
          if (++stage == nstages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (issuer) {
          if (CTAS == 2) umma_commit_2sm(&tfull_bar[acc]);
          else umma_commit(&tfull_bar[acc]);
        }
        if (issuer && tile == unit) GT_STAMP(5);
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------ epilogue (warps 2..9) ---------------------
    const int quad = warp & 3;          // TMEM lane quadrant this warp may access
    const int group = (warp - 2) >> 2;  // 0 or 1: alternate column chunks
    const int row_in_tile = quad * 32 + lane;
    const bool leader = (quad == 2 && lane == 0);  // first thread of the group (warps 2 and 6)
    uint8_t* stage_buf = smem_c + group * 16384;
    // BN = 64, 16-bit output, no residual: a tile is ONE 64-column chunk, which would leave the second warp group idle.
    // The two groups take alternate TILES instead (group g owns the tiles accumulated in TMEM buffer g): both staging
    // buffers and both sets of four warps work, and the epilogue of tile t + 1 overlaps that of tile t.
    const bool alt_tiles = !LN && !GN && BN == 64 && !p.out_f32 && p.tma_store && p.res == nullptr;
    const int cgroup = alt_tiles ? 0 : group;   // chunk-column group
    uint8_t* my_row = stage_buf + row_in_tile * 128;
    const int sw = row_in_tile & 7;
    // columns per chunk: one 128-byte row segment of the output type
    const int chunk_cols = p.out_f32 ? 32 : 64;
    const int n_chunks = BN / chunk_cols;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool store_pending = false;
    const uint32_t tempty_leader0 = (CTAS == 2) ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;
    const uint32_t tempty_leader1 = (CTAS == 2) ? mapa_u32(smem_u32(&tempty_bar[1]), 0) : 0u;
    // ---- residual prefetch (plain GEMM, 16-bit output): each warp TMA-loads the residual of its
    // next (tile, chunk) into its private slice while it is still busy with the current one, so
    // the DRAM latency of the residual never sits on the epilogue's critical path
    uint8_t* rslice = res_slices + (warp - 2) * 4096;
    uint64_t* rbar = &rbar_base[warp - 2];
    uint32_t rphase = 0;
    auto issue_res = [&](int tile_, int ch_) {   // lane 0 only
      const int mu_ = tile_ / p.n_tiles;
      const int nt_ = tile_ - mu_ * p.n_tiles;
      const int mt_ = mu_ * CTAS + static_cast<int>(cta_rank);
      mbar_arrive_expect_tx(rbar, 4096);
      tma_load_4d(rslice, &tmR, rbar, nt_ * BN + ch_ * 64, mt_ * BLOCK_M + quad * 32, 0, 0);
    };
    auto chunk_exists = [&](int tile_, int ch_) {
      if (tile_ >= total_tiles || ch_ >= BN / 64) return false;
      const int nt_ = tile_ % p.n_tiles;
      return nt_ * BN + ch_ * 64 < p.N;
    };
    int ln_tile_count = 0;
    int gn_par = 0;
    if (p.res_tma && lane == 0 && chunk_exists(unit, group)) issue_res(unit, group);
    for (int tile = unit; tile < total_tiles; tile += num_units) {
      const int mu = (p.n_tiles == 1) ? tile : tile / p.n_tiles;
      const int nt = tile - mu * p.n_tiles;
      const int mt = mu * CTAS + static_cast<int>(cta_rank);
      const long long m = static_cast<long long>(mt) * p.tile_rows + row_in_tile;
      const bool row_ok = (row_in_tile < p.tile_rows) && (m < p.M);
      const int n0 = nt * BN;
      int img = 0, h0 = 0;
      if (!p.plain) {
        if (p.nb == 1) {
          img = mt / p.tiles_per_img;
          h0 = (mt - img * p.tiles_per_img) * p.th;
        } else {
          img = mt * p.nb;
        }
      }

      const h16* res_row = nullptr;
      if (p.res != nullptr && row_ok) {
        const long long rr = (p.res_rows > 0) ? (m % p.res_rows) : m;
        res_row = p.res + rr * p.ldr;
      }
      // LayerNorm mode: this thread's 128 residual values are fetched while the MMAs of the tile still run
      uint4 rres[LN ? 16 : 1];
      const bool have_res = LN && res_row != nullptr;
      if constexpr (LN) {
        if (have_res) {
#pragma unroll
          for (int cc = 0; cc < 2; ++cc)
#pragma unroll
            for (int j = 0; j < 8; ++j)
              rres[cc * 8 + j] = __ldg(reinterpret_cast<const uint4*>(res_row + n0 + group * 64 + cc * 128) + j);
        }
      }

      // plain epilogue: this warp's bias values for the whole tile, lane-distributed -- 32-column slot s of the warp's
      // columns lives in bl<s> (16-bit output: chunk group + 2i = slots 2i, 2i + 1; fp32 output: chunk group + 2i = slot i)
      // (scalars, selected with ternaries below: as an indexed array they were put in local memory and fetched with LDL)
      float bl0 = 0.0f, bl1 = 0.0f, bl2 = 0.0f, bl3 = 0.0f;
      uint4 rcur[4];
      const bool res_pre = !LN && !GN && p.out_f32 && p.tma_store && res_row != nullptr;
      auto load_res = [&](int ch_, uint4 (&dst)[4]) {   // fp32 output: residual of one 32-column chunk of this thread's row
        const int nn = n0 + ch_ * 32;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = (ch_ < BN / 32 && nn < p.N) ? __ldg(reinterpret_cast<const uint4*>(res_row + nn) + j) : make_uint4(0, 0, 0, 0);
      };
      if constexpr (!LN && !GN) {
        auto slot_bias = [&](int s_) {
          // slot -> first column: 16-bit: chunk (cgroup + 2*(s_/2)) * 64 + (s_%2) * 32; fp32: chunk (cgroup + 2*s_) * 32
          const int col = p.out_f32 ? (cgroup + 2 * s_) * 32 : (cgroup + 2 * (s_ >> 1)) * 64 + (s_ & 1) * 32;
          const int nn = n0 + col + lane;
          return (p.bias != nullptr && col < BN && nn < p.N) ? __ldg(p.bias + nn) : 0.0f;
        };
        bl0 = slot_bias(0);
        bl1 = slot_bias(1);
        if (BN >= 192) { bl2 = slot_bias(2); bl3 = slot_bias(3); }
        if (res_pre) load_res(group, rcur);
      }
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (tile == unit && warp == 2 && lane == 0) GT_STAMP(6);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * BN);

      if constexpr (GN) {
        // ---------------- GroupNorm folded into the store ----------------
        // The 128 rows of the tile are 128 / gn_hw whole samples and a 64-column chunk holds 64 / cpg whole
        // groups, so the statistics of every (sample, group) of a chunk live in ONE epilogue group: per-thread
        // sums over the group's columns, a segmented warp shuffle over the sample's rows and, for 64-pixel
        // samples, one exchange between the two warps that share the sample.  Fixed order: bit-reproducible.
        const int cpg = p.gn_cpg, hw = p.gn_hw;
        const float inv_cnt = 1.0f / static_cast<float>(hw * cpg);
        const int pair_bar = 3 + group * 2 + (quad >> 1);
#pragma unroll 1
        for (int ch = group; ch < n_chunks; ch += 2) {
          const int c0 = ch * 64;
          const int n = n0 + c0;
          if (n >= p.N) break;
          uint4 rres[8];
          if (res_row != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) rres[j] = __ldg(reinterpret_cast<const uint4*>(res_row + n) + j);
          }
          uint32_t v[64];
          __syncwarp();
          {
            uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
            uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[32]);
            tmem_ld_32x32(taddr + c0, v0);
            tmem_ld_32x32(taddr + c0 + 32, v1);
          }
          tmem_ld_wait();
          // Per-thread sums over fixed 8-column UNITS (static register indices: two instructions per element), folded
          // into groups of cpg = 8 / 16 / 32 / 64 columns afterwards: after the fold, slot u holds the totals of the
          // group that contains unit u, so the normalisation below indexes its mean / rstd statically as well.  (Selecting
          // the group of every column at run time cost ~20 predicated instructions per element.)
          float gs[8], gq[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) gs[u] = gq[u] = 0.0f;
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            const float x = __uint_as_float(v[j]);
            gs[j >> 3] += x;
            gq[j >> 3] = fmaf(x, x, gq[j >> 3]);
          }
          if (cpg >= 16) {
#pragma unroll
            for (int u = 0; u < 8; u += 2) { gs[u] += gs[u + 1]; gq[u] += gq[u + 1]; }
          }
          if (cpg >= 32) {
#pragma unroll
            for (int u = 0; u < 8; u += 4) { gs[u] += gs[u + 2]; gq[u] += gq[u + 2]; }
          }
          if (cpg >= 64) { gs[0] += gs[4]; gq[0] += gq[4]; }
          // group totals over the sample's rows: only the leading slot of every group is reduced
          const int seg = hw < 32 ? hw : 32;
          const int ustep = cpg >> 3;   // 1, 2, 4 or 8 slots per group
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if ((u & (ustep - 1)) == 0) {
              for (int o = 1; o < seg; o <<= 1) {
                gs[u] += __shfl_xor_sync(0xffffffffu, gs[u], o);
                gq[u] += __shfl_xor_sync(0xffffffffu, gq[u], o);
              }
            }
          }
          if (hw == 64) {   // the sample spans this warp and its neighbour (quad ^ 1) of the same epilogue group
            float* mine = gn_x + ((gn_par * 8 + (warp - 2)) * 16);
            float* other = gn_x + ((gn_par * 8 + ((warp - 2) ^ 1)) * 16);
            if (lane == 0) {
#pragma unroll
              for (int u = 0; u < 8; ++u) { mine[2 * u] = gs[u]; mine[2 * u + 1] = gq[u]; }
            }
            named_bar_sync(pair_bar, 64);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              // fold in quad order so that both warps compute bit-identical totals
              const float a = (quad & 1) ? other[2 * u] : gs[u], b = (quad & 1) ? gs[u] : other[2 * u];
              const float c = (quad & 1) ? other[2 * u + 1] : gq[u], d = (quad & 1) ? gq[u] : other[2 * u + 1];
              gs[u] = a + b;
              gq[u] = c + d;
            }
            gn_par ^= 1;
          }
          // every slot takes the totals of its group's leading slot
          if (cpg >= 64) { gs[4] = gs[0]; gq[4] = gq[0]; }
          if (cpg >= 32) {
#pragma unroll
            for (int u = 0; u < 8; u += 4) { gs[u + 2] = gs[u]; gq[u + 2] = gq[u]; }
          }
          if (cpg >= 16) {
#pragma unroll
            for (int u = 0; u < 8; u += 2) { gs[u + 1] = gs[u]; gq[u + 1] = gq[u]; }
          }
          float gmean[8], grstd[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            gmean[u] = gs[u] * inv_cnt;
            grstd[u] = rsqrtf(fmaxf(gq[u] * inv_cnt - gmean[u] * gmean[u], 0.0f) + 1e-5f);
          }
          if (store_pending) {
            if (p.plain) {
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            } else {
              if (leader) tma_store_wait_read<0>();
              named_bar_sync(1 + group, 128);
            }
          }
#pragma unroll
          for (int j8 = 0; j8 < 8; ++j8) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; e += 4) {
              const int j = j8 * 8 + e;
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.gn_gamma + n + j));
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.gn_beta + n + j));
              const float gam[4] = {g4.x, g4.y, g4.z, g4.w}, bet[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int t = 0; t < 4; ++t)
                f[e + t] = (__uint_as_float(v[j + t]) - gmean[j8]) * grstd[j8] * gam[t] + bet[t];
            }
            if (res_row != nullptr) {
              const uint4 r4 = rres[j8];
              float2 t;
              t = unpack_h2(r4.x); f[0] += t.x; f[1] += t.y;
              t = unpack_h2(r4.y); f[2] += t.x; f[3] += t.y;
              t = unpack_h2(r4.z); f[4] += t.x; f[5] += t.y;
              t = unpack_h2(r4.w); f[6] += t.x; f[7] += t.y;
            }
            uint4 q;
            if (p.act == ACT_RELU) {
              q.x = pack_h2_relu(f[0], f[1]); q.y = pack_h2_relu(f[2], f[3]); q.z = pack_h2_relu(f[4], f[5]); q.w = pack_h2_relu(f[6], f[7]);
            } else {
              q.x = pack_h2(f[0], f[1]); q.y = pack_h2(f[2], f[3]); q.z = pack_h2(f[4], f[5]); q.w = pack_h2(f[6], f[7]);
            }
            *reinterpret_cast<uint4*>(my_row + ((j8 ^ sw) << 4)) = q;
          }
          fence_proxy_async();
          if (p.plain) {
            __syncwarp();
            if (lane == 0) {
              tma_store_4d(&tmC, stage_buf + quad * 4096, n, mt * BLOCK_M + quad * 32, 0, 0);
              tma_store_commit();
            }
          } else {
            named_bar_sync(1 + group, 128);
            if (leader) {
              tma_store_4d(&tmC, stage_buf, n, 0, h0, img);
              tma_store_commit();
            }
          }
          store_pending = true;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if constexpr (LN) {
        // ---------------- LayerNorm folded into the store ----------------
        // This thread owns output row `row_in_tile`; its warp group covers column chunks group, group + 2 of
        // this CTA's 256 columns.  Pass 1 adds bias / residual / activation, writes the values back to TMEM
        // and accumulates (sum, sum of squares) in fp32; the partials of the 2 groups x ncl CTAs are pushed
        // into every CTA of the cluster (DSMEM) and folded in a fixed order -> mean, rstd (one exchange).
        // Pass 2 normalises, applies gamma / beta (+ positional table) and stores through TMA.
        float2* ln_buf = reinterpret_cast<float2*>(smem_a + nstages * A_STAGE_BYTES);   // [2 parity][ncl][2 group][128]
        const int par = ln_tile_count & 1;
        auto ln_slot = [&](int src, int grp) { return ln_buf + (((par * ncl + src) * 2 + grp) * BLOCK_M); };
        // pass 1
        float sum = 0.0f, sq = 0.0f;
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = group * 64 + cc * 128;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(taddr + c0 + half * 32, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (have_res) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 r4 = rres[(cc * 2 + half) * 4 + j];
                float2 t;
                t = unpack_h2(r4.x); f[8 * j] += t.x; f[8 * j + 1] += t.y;
                t = unpack_h2(r4.y); f[8 * j + 2] += t.x; f[8 * j + 3] += t.y;
                t = unpack_h2(r4.z); f[8 * j + 4] += t.x; f[8 * j + 5] += t.y;
                t = unpack_h2(r4.w); f[8 * j + 6] += t.x; f[8 * j + 7] += t.y;
              }
            }
            epilogue_math(f, p, n0 + c0 + half * 32, nullptr);   // bias + activation
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              sum += f[j];
              sq = fmaf(f[j], f[j], sq);
              v[j] = __float_as_uint(f[j]);
            }
            tmem_st_32x32(taddr + c0 + half * 32, v);
          }
        }
        tmem_st_wait();
        // exchange (sum, sumsq)
        float tsum = 0.0f, tsq = 0.0f;
        if (p.ln_xchg) {
          // independent CTAs: partials through global memory (L2), completion through a monotonic counter per
          // (row block, lane quadrant) that the 2 x n_tiles warps owning these 32 rows increment once per launch
          const int nex = p.n_tiles;
          float2* ws = reinterpret_cast<float2*>(p.ln_ws);
          __stcg(ws + ((static_cast<size_t>(mu) * nex + nt) * 2 + group) * BLOCK_M + row_in_tile, make_float2(sum, sq));
          __threadfence();
          __syncwarp();
          if (lane == 0) {
            int* cnt = p.ln_cnt + mu * 4 + quad;
            const int per_launch = 2 * nex;
            const int old = atomicAdd(cnt, 1);
            const int target = (old / per_launch + 1) * per_launch;
            uint32_t spins = 0;
            while (true) {
              int cur;
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cur) : "l"(cnt) : "memory");
              if (cur >= target) break;
              if (++spins > (1u << 16)) {
                __nanosleep(128);
                if (spins > (1u << 16) + (1u << 24)) __trap();
              }
            }
          }
          __syncwarp();
          for (int r = 0; r < nex; ++r) {
            const float2 a0 = __ldcg(ws + ((static_cast<size_t>(mu) * nex + r) * 2 + 0) * BLOCK_M + row_in_tile);
            const float2 a1 = __ldcg(ws + ((static_cast<size_t>(mu) * nex + r) * 2 + 1) * BLOCK_M + row_in_tile);
            tsum += a0.x + a1.x;
            tsq += a0.y + a1.y;
          }
        } else {
          float2* mine = ln_slot(static_cast<int>(ln_rank), group) + row_in_tile;
          if (ncl == 1) {
            *mine = make_float2(sum, sq);
            mbar_arrive(&ln_bar[0]);
            mbar_wait(&ln_bar[0], static_cast<uint32_t>(par));
          } else {
            const uint32_t a_data = smem_u32(mine), a_bar = smem_u32(&ln_bar[0]);
            for (int r = 0; r < ncl; ++r) {
              const uint32_t ra = mapa_u32(a_data, static_cast<uint32_t>(r));
              asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(ra), "f"(sum), "f"(sq) : "memory");
              mbar_arrive_release_cluster(mapa_u32(a_bar, static_cast<uint32_t>(r)));
            }
            mbar_wait_acquire_cluster(&ln_bar[0], static_cast<uint32_t>(par));
          }
          for (int r = 0; r < ncl; ++r) {
            const float2 a0 = ln_slot(r, 0)[row_in_tile], a1 = ln_slot(r, 1)[row_in_tile];
            tsum += a0.x + a1.x;
            tsq += a0.y + a1.y;
          }
        }
        const float inv_n = 1.0f / static_cast<float>(p.N);
        const float mean = tsum * inv_n;
        const float rstd = rsqrtf(fmaxf(tsq * inv_n - mean * mean, 0.0f) + p.ln_eps);
        // pass 2
        const float* pe_row = (p.ln_pe != nullptr && row_ok) ? p.ln_pe + static_cast<long long>(m % p.ln_pe_rows) * p.N : nullptr;
#pragma unroll 1
        for (int c0 = group * 64; c0 < BN; c0 += 128) {
          const int n = n0 + c0;
          if (store_pending) {
            if (lane == 0) tma_store_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(taddr + c0 + half * 32, v);
            tmem_ld_wait();
            float f[32];
            const int nn = n + half * 32;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(p.ln_gamma + nn + j));
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.ln_beta + nn + j));
              f[j] = (__uint_as_float(v[j]) - mean) * rstd * g4.x + b4.x;
              f[j + 1] = (__uint_as_float(v[j + 1]) - mean) * rstd * g4.y + b4.y;
              f[j + 2] = (__uint_as_float(v[j + 2]) - mean) * rstd * g4.z + b4.z;
              f[j + 3] = (__uint_as_float(v[j + 3]) - mean) * rstd * g4.w + b4.w;
              if (pe_row != nullptr) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(pe_row + nn + j));
                f[j] += p4.x; f[j + 1] += p4.y; f[j + 2] += p4.z; f[j + 3] += p4.w;
              }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 q;
              q.x = pack_h2(f[8 * j], f[8 * j + 1]);
              q.y = pack_h2(f[8 * j + 2], f[8 * j + 3]);
              q.z = pack_h2(f[8 * j + 4], f[8 * j + 5]);
              q.w = pack_h2(f[8 * j + 6], f[8 * j + 7]);
              *reinterpret_cast<uint4*>(my_row + (((half * 4 + j) ^ sw) << 4)) = q;
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmC, stage_buf + quad * 4096, n, mt * BLOCK_M + quad * 32, 0, 0);
            tma_store_commit();
          }
          store_pending = true;
        }
        ++ln_tile_count;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (!LN && !GN && p.tma_store) {
#pragma unroll 1
        for (int ch = (alt_tiles && acc != group) ? n_chunks : cgroup; ch < n_chunks; ch += 2) {
          const int c0 = ch * chunk_cols;
          const int n = n0 + c0;
          if (n >= p.N) break;  // uniform across the group
          // the staging buffer is free once the previous store from it has been read out
          if (store_pending) {
            if (p.plain) {   // per-warp stores: only this warp's previous store has to have drained
              if (lane == 0) tma_store_wait_read<0>();
              __syncwarp();
            } else {
              if (leader) tma_store_wait_read<0>();
              named_bar_sync(1 + group, 128);
            }
          }
          if (p.out_f32) {
            uint32_t v[32];
            __syncwarp();
            tmem_ld_32x32(taddr + c0, v);
            tmem_ld_wait();
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            if (res_pre) {     // the next chunk's residual is in flight while this one is processed
              uint4 rnext[4];
              load_res(ch + 2, rnext);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float2 t;
                t = unpack_h2(rcur[j].x); f[8 * j] += t.x; f[8 * j + 1] += t.y;
                t = unpack_h2(rcur[j].y); f[8 * j + 2] += t.x; f[8 * j + 3] += t.y;
                t = unpack_h2(rcur[j].z); f[8 * j + 4] += t.x; f[8 * j + 5] += t.y;
                t = unpack_h2(rcur[j].w); f[8 * j + 6] += t.x; f[8 * j + 7] += t.y;
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) rcur[j] = rnext[j];
            }
            {
              const int bi = (ch - cgroup) >> 1;
              const float bsel = (bi == 0) ? bl0 : (bi == 1) ? bl1 : (bi == 2) ? bl2 : bl3;
              epilogue_math_b(f, p, n, res_pre ? nullptr : res_row, bsel);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j)
              *reinterpret_cast<float4*>(my_row + ((j ^ sw) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          } else {
            if (p.res_tma) {   // this chunk's residual slice has landed
              mbar_wait(rbar, rphase);
              rphase ^= 1;
            }
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              uint32_t v[32];
              __syncwarp();
              if (!(p.dbg & 4)) tmem_ld_32x32(taddr + c0 + half * 32, v);
              else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
              }
              uint4 rv[4];
              if (p.res_tma) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  rv[j] = *reinterpret_cast<const uint4*>(rslice + lane * 128 + (((half * 4 + j) ^ (lane & 7)) << 4));
                if (half == 1) {
                  // every lane has read its row: the slice can take the next (tile, chunk)
                  __syncwarp();
                  if (lane == 0) {
                    if (chunk_exists(tile, ch + 2)) issue_res(tile, ch + 2);
                    else if (chunk_exists(tile + num_units, group)) issue_res(tile + num_units, group);
                  }
                }
              }
              tmem_ld_wait();
              float f[32];
#pragma unroll
              for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
              if (p.res_tma) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  float2 t;
                  t = unpack_h2(rv[j].x); f[8 * j] += t.x; f[8 * j + 1] += t.y;
                  t = unpack_h2(rv[j].y); f[8 * j + 2] += t.x; f[8 * j + 3] += t.y;
                  t = unpack_h2(rv[j].z); f[8 * j + 4] += t.x; f[8 * j + 5] += t.y;
                  t = unpack_h2(rv[j].w); f[8 * j + 6] += t.x; f[8 * j + 7] += t.y;
                }
              }
              if (!(p.dbg & 2)) {
                const int bi = (ch - cgroup) + half;      // slot 2 * ((ch - cgroup) / 2) + half
                const float bsel = (bi == 0) ? bl0 : (bi == 1) ? bl1 : (bi == 2) ? bl2 : bl3;
                epilogue_math_b(f, p, n + half * 32, p.res_tma ? nullptr : res_row, bsel, true);
              }
              if (p.act == ACT_RELU) {   // ReLU, saturation, rounding and packing in one F2FP per pair
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 q;
                  q.x = pack_h2_relu(f[8 * j], f[8 * j + 1]);
                  q.y = pack_h2_relu(f[8 * j + 2], f[8 * j + 3]);
                  q.z = pack_h2_relu(f[8 * j + 4], f[8 * j + 5]);
                  q.w = pack_h2_relu(f[8 * j + 6], f[8 * j + 7]);
                  *reinterpret_cast<uint4*>(my_row + (((half * 4 + j) ^ sw) << 4)) = q;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  uint4 q;
                  q.x = pack_h2(f[8 * j], f[8 * j + 1]);
                  q.y = pack_h2(f[8 * j + 2], f[8 * j + 3]);
                  q.z = pack_h2(f[8 * j + 4], f[8 * j + 5]);
                  q.w = pack_h2(f[8 * j + 6], f[8 * j + 7]);
                  *reinterpret_cast<uint4*>(my_row + (((half * 4 + j) ^ sw) << 4)) = q;
                }
              }
            }
          }
          fence_proxy_async();  // make the st.shared visible to the TMA (async proxy)
          if (p.plain) {
            // plain GEMM: the output box is 32 rows, so every warp stores its own quarter of the
            // tile as soon as it is staged -- no cross-warp barrier, the eight warps run decoupled
            __syncwarp();
            if (lane == 0 && !(p.dbg & 1)) {
              tma_store_4d(&tmC, stage_buf + quad * 4096, n, mt * BLOCK_M + quad * 32, 0, 0);
              tma_store_commit();
            }
          } else {
            named_bar_sync(1 + group, 128);
            if (leader) {
              tma_store_4d(&tmC, stage_buf, n, 0, h0, img);
              tma_store_commit();
            }
          }
          store_pending = true;
        }
        // every TMEM read of this warp for this accumulator has completed (wait::ld above)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
          else mbar_arrive(&tempty_bar[acc]);
        }
      } else if (!LN && !GN) {
        // direct global stores (validation path, ROBOVLN_EPILOGUE=direct): 32-column chunks
#pragma unroll 1
        for (int c0 = group * 32; c0 < BN; c0 += 64) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_32x32(taddr + c0, v);
          tmem_ld_wait();
          const int n = n0 + c0;
          if (row_ok && n < p.N) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
            epilogue_math(f, p, n, res_row);
            if (p.out_f32) {
              float* o = reinterpret_cast<float*>(p.out) + m * p.ldc + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (n + j < p.N) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
              }
            } else {
              h16* o = reinterpret_cast<h16*>(p.out) + m * p.ldc + n;
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                if (n + j < p.N) {
                  uint4 q;
                  q.x = pack_h2(f[j], f[j + 1]);
                  q.y = pack_h2(f[j + 2], f[j + 3]);
                  q.z = pack_h2(f[j + 4], f[j + 5]);
                  q.w = pack_h2(f[j + 6], f[j + 7]);
                  *reinterpret_cast<uint4*>(o + j) = q;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CTAS == 2) mbar_arrive_cluster(acc ? tempty_leader1 : tempty_leader0);
          else mbar_arrive(&tempty_bar[acc]);
        }
      }
      if (tile == unit && warp == 2 && lane == 0) GT_STAMP(7);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (warp == 2 && lane == 0) GT_STAMP(8);
    // staging smem must stay valid until the last bulk store has read it
    if (store_pending && (p.plain ? (lane == 0) : leader)) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (CTAS == 2 || ncl > 1) cluster_sync_all();   // the peer may still be using this CTA's barriers / TMEM half / smem
  if (warp == 1) {
    tc_fence_after();
    if (CTAS == 2) tmem_dealloc_2sm<C::TMEM_COLS>(tmem_base);
    else tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(ptr);
  });
  RVB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  return fn;
}

void encode_map(CUtensorMap* map, CUtensorMapDataType dt, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* estr) {
  cuuint64_t gd[5];
  cuuint64_t gs[4];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = estr[i];
  }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = strides_bytes[i];
  CUresult r = encode_tiled_fn()(map, dt, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gd, gs, bx, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    std::string msg = "cuTensorMapEncodeTiled failed (" + std::to_string(static_cast<int>(r)) + ") dims=";
    for (int i = 0; i < rank; ++i) msg += std::to_string(dims[i]) + (i + 1 < rank ? "x" : "");
    msg += " box=";
    for (int i = 0; i < rank; ++i) msg += std::to_string(box[i]) + (i + 1 < rank ? "x" : "");
    msg += " strides=";
    for (int i = 0; i + 1 < rank; ++i) msg += std::to_string(strides_bytes[i]) + (i + 2 < rank ? "," : "");
    throw Error(static_cast<int>(r), msg);
  }
}

constexpr CUtensorMapDataType kH16Type = RVB_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;

template <int BN, int CTAS, int EPI = 0>
void launch_bn(const GemmTcPlan& plan, cudaStream_t stream) {
  using C = Cfg<BN, CTAS>;
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    RVB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CTAS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
  }
  cudaLaunchConfig_t cfg;
  std::memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(plan.grid, 1, 1);
  cfg.blockDim = dim3(NUM_THREADS, 1, 1);
  cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int na = 0;
  if (g_launch_prio != 0) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = g_launch_prio;
    ++na;
  }
  const int cluster_x = (CTAS == 2) ? 2 : ((EPI == 1 && !plan.p.ln_xchg) ? plan.p.n_tiles : 1);
  if (cluster_x > 1) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = cluster_x;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (use_pdl()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
#if defined(RVB_GEMM_STAMPS) && RVB_GEMM_STAMPS
  // ROBOVLN_GEMM_TIMES=<file>: per-CTA phase stamps of every launch (diagnostic builds only; synchronises the stream)
  static const char* tenv = std::getenv("ROBOVLN_GEMM_TIMES");
  if (tenv != nullptr) {
    GemmTcParams pp = plan.p;
    const size_t tbytes = static_cast<size_t>(plan.grid) * 16 * sizeof(long long);
    RVB_CUDA(cudaMalloc(&pp.times, tbytes));
    RVB_CUDA(cudaMemsetAsync(pp.times, 0, tbytes, stream));
    RVB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CTAS, EPI>, plan.tmA, plan.tmB, plan.tmC, plan.tmR, pp));
    RVB_CUDA(cudaStreamSynchronize(stream));
    std::vector<long long> host(static_cast<size_t>(plan.grid) * 16);
    RVB_CUDA(cudaMemcpy(host.data(), pp.times, tbytes, cudaMemcpyDeviceToHost));
    RVB_CUDA(cudaFree(pp.times));
    if (FILE* f = std::fopen(tenv, "a")) {
      std::fprintf(f, "# gemm BN=%d CTAS=%d EPI=%d M=%d N=%d num_kb=%d grid=%d tiles=%d act=%d out_f32=%d res=%d\n", BN, CTAS, EPI, plan.p.M,
                   plan.p.N, plan.p.num_kb, plan.grid, plan.p.m_tiles * plan.p.n_tiles, plan.p.act, plan.p.out_f32, plan.p.res != nullptr);
      for (int c = 0; c < plan.grid; ++c) {
        std::fprintf(f, "%d", c);
        for (int i = 0; i < 16; ++i) std::fprintf(f, ",%lld", host[static_cast<size_t>(c) * 16 + i]);
        std::fprintf(f, "\n");
      }
      std::fclose(f);
    }
    return;
  }
#endif
  RVB_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CTAS, EPI>, plan.tmA, plan.tmB, plan.tmC, plan.tmR, plan.p));
}

bool use_pair_mma() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_PAIR_MMA");
    v = (e != nullptr && std::strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}

bool use_res_tma() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_RES_TMA");
    v = (e != nullptr && std::strcmp(e, "0") == 0) ? 0 : 1;
  }
  return v == 1;
}

bool use_direct_epilogue() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_EPILOGUE");
    v = (e != nullptr && std::strcmp(e, "direct") == 0) ? 1 : 0;
  }
  return v == 1;
}

}  // namespace

thread_local int g_pdl_override = -1;
thread_local int g_launch_prio = 0;

void tma_encode_2d_h16(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_cols,
                       uint32_t box_rows) {
  const uint64_t dims[2] = {cols, rows};
  const uint64_t strides[1] = {pitch_bytes};
  const uint32_t box[2] = {box_cols, box_rows};
  const uint32_t ones[2] = {1, 1};
  encode_map(map, kH16Type, base, 2, dims, strides, box, ones);
}

void tma_encode_nd_h16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box) {
  const uint32_t ones[5] = {1, 1, 1, 1, 1};
  encode_map(map, kH16Type, base, rank, dims, strides_bytes, box, ones);
}

bool use_pdl() {
  if (g_pdl_override >= 0) return g_pdl_override == 1;
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_PDL");
    v = (e != nullptr && std::strcmp(e, "1") == 0) ? 1 : 0;   // measured slower on B200: off by default
  }
  return v == 1;
}

int device_sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    RVB_CUDA(cudaGetDevice(&dev));
    RVB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  return sms;
}

bool use_simt_gemm() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("ROBOVLN_GEMM");
    v = (e != nullptr && std::strcmp(e, "simt") == 0) ? 1 : 0;
  }
  return v == 1;
}

void gemm_tc_make_plan(const ConvGemm& g, GemmTcPlan* plan, int force_bn) {
  RVB_CHECK(g.in != nullptr && g.w != nullptr && g.out != nullptr, "gemm: null operand");
  RVB_CHECK(g.Cin % 8 == 0 && (g.window == 2 || g.in_pitch % 8 == 0), "gemm: Cin / pitch must be multiples of 8");
  RVB_CHECK(g.window || g.in_pitch >= g.Cin, "gemm: pixel pitch smaller than Cin");
  RVB_CHECK((reinterpret_cast<uintptr_t>(g.in) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.w) & 15) == 0,
            "gemm: operands must be 16-byte aligned");
  RVB_CHECK(g.Cout % 8 == 0, "gemm: Cout must be a multiple of 8");
  RVB_CHECK(g.ldc % (g.out_f32 ? 4 : 8) == 0 && (reinterpret_cast<uintptr_t>(g.out) & 15) == 0,
            "gemm: output pitch/alignment");
  if (g.res != nullptr)
    RVB_CHECK(g.ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(g.res) & 15) == 0, "gemm: residual pitch/alignment");
  if (g.bias != nullptr) RVB_CHECK((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "gemm: bias alignment");

  plan->desc = g;
  GemmTcParams& p = plan->p;
  std::memset(&p, 0, sizeof(p));
  const int Ho = g.Ho(), Wo = g.Wo();
  const long long M = g.M();
  RVB_CHECK(M > 0 && M < (1ll << 31), "gemm: M out of range");
  p.M = static_cast<int>(M);
  p.N = g.Cout;
  p.Cin = g.Cin;
  p.KW = g.KW;
  p.cblocks = (g.Cin + BLOCK_K - 1) / BLOCK_K;
  p.num_kb = g.KH * g.KW * p.cblocks;
  p.stride = g.stride;
  p.pad = g.pad;
  p.plain = g.plain() ? 1 : 0;
  p.bias = g.bias;
  p.res = g.res;
  p.ldr = g.ldr;
  p.res_rows = g.res_rows;
  p.act = g.act;
  p.out = g.out;
  p.ldc = g.ldc;
  p.out_f32 = g.out_f32;
  p.tma_store = use_direct_epilogue() ? 0 : 1;
  const bool gn = g.gn_gamma != nullptr;
  if (gn) {
    RVB_CHECK(!g.out_f32 && g.gn_beta != nullptr && g.gn_groups > 0 && g.Cout % g.gn_groups == 0 && p.tma_store && g.res_rows == 0 &&
                  g.bias == nullptr && (g.act == ACT_NONE || g.act == ACT_RELU), "gemm: GroupNorm epilogue: unsupported combination");
    const int cpg = g.Cout / g.gn_groups;
    RVB_CHECK((cpg == 8 || cpg == 16 || cpg == 32 || cpg == 64) && (g.gn_hw == 16 || g.gn_hw == 64) && g.gn_hw == Ho * Wo,
              "gemm: GroupNorm epilogue needs 8..64 channels per group and 16- or 64-pixel samples");
    p.gn_hw = g.gn_hw; p.gn_cpg = cpg; p.gn_gamma = g.gn_gamma; p.gn_beta = g.gn_beta;
    if (force_bn == 0) force_bn = (g.Cout >= 128 && g.Cout % 128 == 0) ? 128 : 64;
  }
  const bool ln = g.ln_gamma != nullptr;
  if (ln) {
    RVB_CHECK(g.plain() && !g.out_f32 && g.ln_beta != nullptr && g.Cout % 256 == 0 && g.Cout <= 768 && p.tma_store,
              "gemm: LayerNorm epilogue needs a plain GEMM with h16 output and Cout in {256, 512, 768}");
    RVB_CHECK((reinterpret_cast<uintptr_t>(g.ln_gamma) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.ln_beta) & 15) == 0 &&
                  (g.ln_pe == nullptr || (reinterpret_cast<uintptr_t>(g.ln_pe) & 15) == 0), "gemm: LayerNorm parameter alignment");
    p.ln = 1; p.ln_eps = g.ln_eps; p.ln_gamma = g.ln_gamma; p.ln_beta = g.ln_beta; p.ln_pe = g.ln_pe;
    p.ln_pe_rows = g.ln_pe_rows > 0 ? g.ln_pe_rows : 1;
    p.ln_xchg = (g.ln_ws != nullptr && g.ln_cnt != nullptr && g.Cout > 256) ? 1 : 0;
    p.ln_ws = g.ln_ws; p.ln_cnt = g.ln_cnt;
    force_bn = 256;   // one full 256-column tile per CTA, 1-CTA form
  }
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = std::getenv("ROBOVLN_EPI_DEBUG");
      dbg = e ? std::atoi(e) : 0;
    }
    p.dbg = dbg;
  }

  const uint32_t ones[4] = {1, 1, 1, 1};
  const uint64_t pitchB = static_cast<uint64_t>(g.in_pitch) * 2;
  if (p.plain) {
    p.tile_rows = BLOCK_M;
    p.th = 1;
    p.nb = 1;
    p.tiles_per_img = 1;
    p.m_tiles = static_cast<int>((M + BLOCK_M - 1) / BLOCK_M);
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(M), 1, 1};
    const uint64_t strides[3] = {pitchB, pitchB * static_cast<uint64_t>(M), pitchB * static_cast<uint64_t>(M)};
    const uint32_t box[4] = {BLOCK_K, BLOCK_M, 1, 1};
    encode_map(&plan->tmA, kH16Type, g.in, 4, dims, strides, box, ones);
  } else {
    RVB_CHECK(Wo <= BLOCK_M, "conv: output width > 128 is not supported by the full-row tiling");
    RVB_CHECK(g.window || Wo * g.stride <= 256, "conv: box width exceeds the TMA limit");
    const int P = Ho * Wo;
    if (P >= BLOCK_M) {
      int th = 1;
      for (int t = 1; t <= Ho; ++t)
        if (Ho % t == 0 && t * Wo <= BLOCK_M && t * g.stride <= 256) th = t;
      p.th = th;
      p.nb = 1;
      p.tiles_per_img = Ho / th;
      p.tile_rows = th * Wo;
      p.m_tiles = g.NB * p.tiles_per_img;
    } else {
      p.th = Ho;
      p.nb = BLOCK_M / P;
      p.tiles_per_img = 1;
      p.tile_rows = p.nb * P;
      p.m_tiles = (g.NB + p.nb - 1) / p.nb;
    }
    if (g.window == 2) {
      // packed stem: element (k, wo, t, n) lives at n*H*rowpair + t*rowpair + wo*(stride px * 8) + k, k < 64
      RVB_CHECK(g.KH == 4 && g.KW == 1 && g.pad == 0 && g.Cin == BLOCK_K && g.win_row_pitch % 8 == 0,
                "packed window conv: bad geometry");
      RVB_CHECK(static_cast<int64_t>(Wo - 1) * g.stride * 8 + BLOCK_K <= g.win_row_pitch, "packed window conv: row too short");
      const uint64_t rowB = static_cast<uint64_t>(g.win_row_pitch) * 2;
      const uint64_t dims[4] = {BLOCK_K, static_cast<uint64_t>(Wo), static_cast<uint64_t>(g.H), static_cast<uint64_t>(g.NB)};
      const uint64_t strides[3] = {static_cast<uint64_t>(g.stride) * 8 * 2, rowB, rowB * g.H};
      const uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(Wo), static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.nb)};
      encode_map(&plan->tmA, kH16Type, g.in, 4, dims, strides, box, ones);
      p.window2 = 1;
    } else if (g.window) {
      // overlapping-window view of a zero-padded NHW8 image: element (k, wo, h, n) lives at
      // n*H*row + h*row + wo*(stride*8) + k; each K block is one filter ROW (KW folded into k).
      RVB_CHECK(g.KW == 1 && g.pad == 0 && g.Cin == BLOCK_K && g.win_row_pitch % 8 == 0, "window conv: bad geometry");
      RVB_CHECK(static_cast<int64_t>(Wo - 1) * g.stride * 8 + BLOCK_K <= g.win_row_pitch, "window conv: row too short");
      const uint64_t rowB = static_cast<uint64_t>(g.win_row_pitch) * 2;
      const uint64_t dims[4] = {BLOCK_K, static_cast<uint64_t>(Wo), static_cast<uint64_t>(g.H),
                                static_cast<uint64_t>(g.NB)};
      const uint64_t strides[3] = {static_cast<uint64_t>(g.stride) * 8 * 2, rowB, rowB * g.H};
      const uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(Wo), static_cast<uint32_t>(p.th * g.stride),
                               static_cast<uint32_t>(p.nb)};
      const uint32_t es[4] = {1, 1, static_cast<uint32_t>(g.stride), 1};
      encode_map(&plan->tmA, kH16Type, g.in, 4, dims, strides, box, es);
    } else {
      const uint64_t dims[4] = {static_cast<uint64_t>(g.Cin), static_cast<uint64_t>(g.W), static_cast<uint64_t>(g.H),
                                static_cast<uint64_t>(g.NB)};
      const uint64_t strides[3] = {pitchB, pitchB * g.W, pitchB * g.W * g.H};
      const uint32_t box[4] = {BLOCK_K, static_cast<uint32_t>(Wo * g.stride), static_cast<uint32_t>(p.th * g.stride),
                               static_cast<uint32_t>(p.nb)};
      const uint32_t es[4] = {1, static_cast<uint32_t>(g.stride), static_cast<uint32_t>(g.stride), 1};
      encode_map(&plan->tmA, kH16Type, g.in, 4, dims, strides, box, es);
    }
  }
  p.a_bytes = static_cast<uint32_t>(p.tile_rows) * BLOCK_K * 2;

  // Tile configuration: minimise (waves x tile work / relative tile efficiency).  The relative
  // efficiencies reflect the operand bytes each SM must pull from L2 per FLOP (128x128: 1/64,
  // 128x256: 1/85, pair 256x256: 1/128 B/FLOP) as measured on B200 (profiles/): the 1-CTA forms
  // are ingress-bound long before the tensor pipe saturates.  The pair form needs enough K to
  // amortise its cluster hand-shakes and is only built for BN = 256.
  const int sms = device_sm_count();
  struct Cand { int bn, ctas; double eff; };
  const Cand cands[4] = {{256, 2, 1.0}, {256, 1, 0.62}, {128, 1, 0.5}, {64, 1, 0.33}};
  const long long Kdepth = static_cast<long long>(g.KH) * g.KW * g.Cin;
  int best_bn = 0, best_ctas = 1;
  double best_cost = 1e300;
  for (const Cand& c : cands) {
    if (force_bn != 0 && (c.bn != (force_bn & 0xffff) || c.ctas != ((force_bn >> 16) ? 2 : 1))) continue;
    if (force_bn == 0) {
      if (c.bn > 64 && g.Cout <= c.bn / 2) continue;          // more than half the tile would be padding
      // measured (profiles/r01_gemm_probe.txt): the pair form wins once there are >= 2 full waves of
      // 256 x 256 tiles and the epilogue is light; below that its coarser tiles lose to quantisation
      const long long pair_units = static_cast<long long>((p.m_tiles + 1) / 2) * ((g.Cout + 255) / 256);
      if (c.ctas == 2 && (Kdepth < 512 || g.Cout < 256 || pair_units < 2 * (sms / 2) || g.act == ACT_GELU ||
                          !use_pair_mma())) continue;
    }
    const int nt = (g.Cout + c.bn - 1) / c.bn;
    const long long units = static_cast<long long>((p.m_tiles + c.ctas - 1) / c.ctas) * nt;
    const long long slots = sms / c.ctas;
    const long long waves = (units + slots - 1) / slots;
    const double cost = static_cast<double>(waves) * c.bn / c.eff;   // time ~ waves x per-SM tile area (128 x BN) / efficiency
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best_bn = c.bn;
      best_ctas = c.ctas;
    }
  }
  RVB_CHECK(best_bn != 0, "gemm: no tile configuration");
  plan->BN = best_bn;
  plan->ctas = best_ctas;
  p.n_tiles = (g.Cout + best_bn - 1) / best_bn;

  const uint64_t Ktot = static_cast<uint64_t>(g.KH) * g.KW * g.Cin;
  const uint64_t bdims[2] = {Ktot, static_cast<uint64_t>(g.Cout)};
  const uint64_t bstrides[1] = {Ktot * 2};
  const uint32_t bbox[2] = {BLOCK_K, static_cast<uint32_t>(best_bn / best_ctas)};
  RVB_CHECK((Ktot * 2) % 16 == 0, "gemm: weight row pitch must be a multiple of 16 bytes");
  encode_map(&plan->tmB, kH16Type, g.w, 2, bdims, bstrides, bbox, ones);

  // output map (TMA store): same row geometry as the A tile, 128-byte column chunks
  {
    const uint64_t esz = g.out_f32 ? 4 : 2;
    const uint32_t ccols = g.out_f32 ? 32 : 64;
    const uint64_t pitchC = static_cast<uint64_t>(g.ldc) * esz;
    const CUtensorMapDataType cdt = g.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : kH16Type;
    if (p.plain) {
      const uint64_t dims[4] = {static_cast<uint64_t>(g.Cout), static_cast<uint64_t>(M), 1, 1};
      const uint64_t strides[3] = {pitchC, pitchC * static_cast<uint64_t>(M), pitchC * static_cast<uint64_t>(M)};
      const uint32_t box[4] = {ccols, 32, 1, 1};   // one epilogue warp's rows (per-warp stores)
      encode_map(&plan->tmC, cdt, g.out, 4, dims, strides, box, ones);
    } else {
      const uint64_t dims[4] = {static_cast<uint64_t>(g.Cout), static_cast<uint64_t>(Wo), static_cast<uint64_t>(Ho),
                                static_cast<uint64_t>(g.NB)};
      const uint64_t strides[3] = {pitchC, pitchC * Wo, pitchC * Wo * Ho};
      const uint32_t box[4] = {ccols, static_cast<uint32_t>(Wo), static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.nb)};
      encode_map(&plan->tmC, cdt, g.out, 4, dims, strides, box, ones);
    }
  }

  // residual prefetch through TMA (plain GEMM, 16-bit output, one residual row per output row)
  p.res_tma = (p.tma_store && p.plain && !g.out_f32 && g.res != nullptr && g.res_rows == 0 && use_res_tma() && !ln && !gn) ? 1 : 0;
  {
    const int b_stage = (best_bn / best_ctas) * BLOCK_K * 2;
    const int max_stages = std::min(8, SMEM_STAGE_BUDGET / (A_STAGE_BYTES + b_stage));
    // the residual slices (32 KiB) live in the B buffers of the stages given up
    p.nstages = p.res_tma ? max_stages - (32768 + b_stage - 1) / b_stage : max_stages;
    if (ln) p.nstages = max_stages - 1;   // the A buffer of the stage given up holds the statistics exchange (12 KB)
    // resident B: a launch with ONE N tile multiplies every M tile by the same weights; when they fit beside at least
    // three A stages they are loaded once per CTA and the ring carries A only.  The 64- and 128-wide convs of the trunk
    // fronts re-fetched 8-16 KB of weights per 16 KB of activations -- a third to a half of their L2 -> SM traffic.
    static int bres_env = -1;
    if (bres_env < 0) {
      const char* e = std::getenv("ROBOVLN_B_RESIDENT");
      bres_env = (e != nullptr && std::strcmp(e, "0") == 0) ? 0 : 1;
    }
    const int b_total = p.num_kb * b_stage;
    const int a_stages_left = (SMEM_STAGE_BUDGET - b_total) / A_STAGE_BYTES;
    p.b_res = (bres_env && best_ctas == 1 && p.n_tiles == 1 && !p.res_tma && !ln && !gn && p.m_tiles > 1 && b_total <= SMEM_STAGE_BUDGET &&
               a_stages_left >= 3) ? 1 : 0;
    if (p.b_res) p.nstages = std::min(max_stages, a_stages_left);   // the barrier arrays hold Cfg::STAGES = max_stages entries
    RVB_CHECK(p.nstages >= 2, "gemm: too few pipeline stages");
  }
  if (p.res_tma) {
    const uint64_t pitchR = static_cast<uint64_t>(g.ldr) * 2;
    const uint64_t dims[4] = {static_cast<uint64_t>(g.Cout), static_cast<uint64_t>(M), 1, 1};
    const uint64_t strides[3] = {pitchR, pitchR * static_cast<uint64_t>(M), pitchR * static_cast<uint64_t>(M)};
    const uint32_t box[4] = {64, 32, 1, 1};
    encode_map(&plan->tmR, kH16Type, g.res, 4, dims, strides, box, ones);
  } else {
    plan->tmR = plan->tmC;
  }

  const long long units = static_cast<long long>((p.m_tiles + best_ctas - 1) / best_ctas) * p.n_tiles;
  // balanced persistent grid: with R = ceil(units / slots) rounds, ceil(units / R) CTAs (or pairs) finish in the same R
  // tile times as a full grid would, and the CTAs not launched do not pay the ~3 us set-up (descriptor fetch, first
  // operand latency) on SMs that the other two streams of the step can use.  g.extra_rounds > 0 goes further and
  // trades kernel latency for SM time: the step is bound by the SM time of its three concurrent streams, not by
  // the length of any one chain
  const long long slots = sms / best_ctas;
  long long rounds = (units + slots - 1) / slots;
  if (!ln && g.extra_rounds > 0) {
    // ROBOVLN_GRID_MIN = n keeps the plain grid where the extra round would leave fewer than n CTAs.  Measured: even
    // the launches of a few dozen tiles are better off with the extra round (3.645 ms/step at 0, 3.663 at 24 and 48,
    // 3.69 at 72, 3.73 at 100), so the default is 0
    static int min_grid = -1;
    if (min_grid < 0) {
      const char* e = std::getenv("ROBOVLN_GRID_MIN");
      min_grid = e != nullptr ? std::atoi(e) : 0;
    }
    const long long r2 = rounds + g.extra_rounds;
    if ((units + r2 - 1) / r2 * best_ctas >= min_grid) rounds = r2;
  }
  plan->grid = static_cast<int>((units + rounds - 1) / rounds) * best_ctas;
  if (ln) plan->grid = std::min(p.m_tiles, sms / p.n_tiles) * p.n_tiles;   // whole clusters / whole row blocks of n_tiles CTAs
  plan->valid = true;
}

void gemm_tc_launch(const GemmTcPlan& plan, cudaStream_t stream) {
  RVB_CHECK(plan.valid, "gemm: plan not built");
  if (use_simt_gemm()) {
    gemm_simt_launch(plan.desc, stream);
    return;
  }
  if (plan.ctas == 2) {
    RVB_CHECK(plan.BN == 256, "gemm: the CTA-pair kernel is built for BN = 256");
    launch_bn<256, 2>(plan, stream);
    return;
  }
  if (plan.p.ln) {
    RVB_CHECK(plan.BN == 256 && plan.ctas == 1, "gemm: the LayerNorm epilogue is built for the 1-CTA BN = 256 form");
    launch_bn<256, 1, 1>(plan, stream);
    return;
  }
  if (plan.p.gn_hw != 0) {
    RVB_CHECK(plan.ctas == 1 && (plan.BN == 64 || plan.BN == 128), "gemm: the GroupNorm epilogue is built for BN = 64 / 128");
    if (plan.BN == 64) launch_bn<64, 1, 2>(plan, stream);
    else launch_bn<128, 1, 2>(plan, stream);
    return;
  }
  switch (plan.BN) {
    case 64: launch_bn<64, 1>(plan, stream); break;
    case 128: launch_bn<128, 1>(plan, stream); break;
    case 256: launch_bn<256, 1>(plan, stream); break;
    default: RVB_CHECK(false, "gemm: bad BN");
  }
}

}  // namespace rvb
