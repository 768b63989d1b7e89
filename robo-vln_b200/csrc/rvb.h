// Internal (C++) declarations shared by the .cu files of librobovln_b200.so.
// The public C ABI is include/robovln_b200.h.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "h16.h"

namespace rvb {


struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define RVB_CHECK(cond, msg)                                                                      \
  do {                                                                                            \
    if (!(cond)) throw ::rvb::Error(-1, std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + (msg)); \
  } while (0)

#define RVB_CUDA(expr)                                                                            \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      throw ::rvb::Error(static_cast<int>(_e), std::string(__FILE__) + ":" + std::to_string(__LINE__) + ": " + \
                                                   #expr + " -> " + cudaGetErrorString(_e));      \
  } while (0)

// Kernel launch with the programmatic-stream-serialization (PDL) attribute; every kernel of the
// library calls griddepcontrol.wait before touching data a predecessor may still be writing.
// ROBOVLN_PDL=0 launches plainly (validation).
bool use_pdl();
extern thread_local int g_pdl_override;
extern thread_local int g_launch_prio;    // 0: none; otherwise cudaLaunchAttributePriority for the launches that follow   // -1: environment default; 0 / 1: forced for the launches that follow
template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (use_pdl()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (g_launch_prio != 0) {
    at[na].id = cudaLaunchAttributePriority;
    at[na].val.priority = g_launch_prio;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  RVB_CUDA(cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...));
}

// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) applies to the CURRENT device only, and one process
// may drive several (the reference trainer keeps hi on cuda:0 and lo on cuda:1, hierarchical_trainer.py:292-296):
// remember the opt-in per device, not per process.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// ---------------------------------------------------------------------------------------
// Convolution-as-GEMM problem:  out[m, n] = act( sum_{tap,c} A(m,tap,c) * W[n, tap*Cin + c] + bias[n] + res[m, n] )
//   A is an NHWC h16 tensor [NB, H, W, Cin] (row pitch in_pitch elements per pixel),
//   m enumerates output pixels (n, ho, wo) row-major; a plain GEMM is the case H=1, W=M, 1x1.
// ---------------------------------------------------------------------------------------
struct ConvGemm {
  // input
  const h16* in = nullptr;
  int NB = 1, H = 1, W = 1, Cin = 0;
  int64_t in_pitch = 0;          // elements between consecutive pixels (>= Cin, multiple of 8)
  // filter
  const h16* w = nullptr;       // [Cout, KH*KW*Cin] K-major, k = (r*KW + s)*Cin + c
  int Cout = 0, KH = 1, KW = 1, stride = 1, pad = 0;
  // epilogue
  const float* bias = nullptr;   // [Cout] or null
  const h16* res = nullptr;     // residual [res_rows, ldr] or null; row = m % res_rows
  int64_t ldr = 0;
  int res_rows = 0;              // 0 -> M
  int act = ACT_NONE;
  void* out = nullptr;           // h16 or f32, [M, ldc]
  int64_t ldc = 0;
  int out_f32 = 0;
  // LayerNorm folded into the store (plain GEMM, h16 output, Cout in {256, 512, 768}):
  //   out = LN(act(acc + bias + res)) * ln_gamma + ln_beta (+ ln_pe[m % ln_pe_rows])
  // The Cout/256 CTAs that hold one 128-row block form a cluster and exchange row statistics over DSMEM.
  const float* ln_gamma = nullptr;
  const float* ln_beta = nullptr;
  float ln_eps = 0.0f;
  const float* ln_pe = nullptr;  // [ln_pe_rows, Cout] fp32 added after the norm, or null
  int ln_pe_rows = 1;
  // Cout > 256: how the Cout/256 CTAs of a row block exchange (sum, sumsq).  Null: thread-block cluster + DSMEM.
  // Non-null: through global memory, no cluster launch -- ln_ws holds m_tiles*n_tiles*256 float2, ln_cnt
  // m_tiles*4 ints that are zero before the first launch (monotonic counters: never reset).
  void* ln_ws = nullptr;
  int* ln_cnt = nullptr;
  // persistent-grid policy: the grid is sized for ceil(units / slots) + extra_rounds tile rounds (gemm_tc_make_plan)
  int extra_rounds = 0;
  // GroupNorm folded into the store, for feature maps small enough that every 128-row M tile holds WHOLE
  // samples (gn_hw = H*W of the output in {16, 64}) and every N tile whole groups:
  //   out = relu?( GN_groups(acc) * gn_gamma + gn_beta + res )     (res: plain h16 residual, optional)
  const float* gn_gamma = nullptr;
  const float* gn_beta = nullptr;
  int gn_groups = 0;
  int gn_hw = 0;
  // "window" mode (RGB stem): `in` is a zero-padded [NB, H, win_row_pitch/8, 8] image, KW is folded
  // into the K dimension (Cin = 64 = 8 pixels x 8 channels per filter row), W is the OUTPUT width.
  // window == 2 (packed stem): `in` is a zero-padded, ROW-PAIR-INTERLEAVED image [NB, H, win_row_pitch/8, 2, 4]
  // (H = padded rows / 2; per pixel column: row 2t then row 2t+1, 4 channels each), so one 64-element window
  // = 8 pixels x 2 rows x 4 channels holds TWO filter rows: KH = 4 row pairs for the 7 filter rows (the 8th
  // has zero weights), k = s*8 + (r%2)*4 + c.  The window steps `stride` (2) pixels in W and ONE row pair in H.
  int window = 0;
  int64_t win_row_pitch = 0;     // elements per padded input row
  // derived
  int Ho() const { return window == 2 ? H - 3 : (H + 2 * pad - KH) / stride + 1; }
  int Wo() const { return window ? W : (W + 2 * pad - KW) / stride + 1; }
  int64_t M() const { return static_cast<int64_t>(NB) * Ho() * Wo(); }
  bool plain() const { return !window && KH == 1 && KW == 1 && stride == 1 && pad == 0; }
};

// Device-side parameter block of the tcgen05 kernel.
struct GemmTcParams {
  int M, N;
  int num_kb;        // taps * cblocks
  int cblocks;       // ceil(Cin / 64)
  int KW;            // taps decompose as r = tap / KW, s = tap % KW
  int Cin;
  int plain;         // 1: A coords (k, m0, 0, 0)
  int tile_rows;     // valid rows per 128-row M tile
  int th, nb;        // output rows / images per tile (conv mode)
  int tiles_per_img; // Ho / th when nb == 1
  int stride, pad;
  int m_tiles, n_tiles;
  uint32_t a_bytes;  // bytes one A box delivers (tile_rows * 128)
  const float* bias;
  const h16* res;
  long long ldr;
  int res_rows;
  int act;
  void* out;
  long long ldc;
  int out_f32;
  int tma_store;     // 1: smem-staged TMA store epilogue, 0: direct global stores (validation)
  int res_tma;       // 1: residual chunks prefetched by TMA into per-warp smem slices
  int window2;       // 1: packed stem (rows advance by one row PAIR per output row: A row coord h0 + tap)
  int gn_hw, gn_cpg; // GroupNorm epilogue: pixels per sample (16 / 64), channels per group (8 .. 64); 0 = off
  const float* gn_gamma;
  const float* gn_beta;
  int ln;            // 1: LayerNorm epilogue; the n_tiles CTAs of a 128-row block are one cluster
  float ln_eps;
  const float* ln_gamma;
  const float* ln_beta;
  const float* ln_pe;
  int ln_pe_rows;
  int ln_xchg;       // 1: statistics exchanged through global memory (ln_ws / ln_cnt), independent CTAs
  void* ln_ws;
  int* ln_cnt;
  int nstages;       // pipeline stages in use (fewer when the residual slices take their place)
  int b_res;         // 1: single N tile whose weights stay in shared memory for the whole launch (the ring carries A only)
  long long* times;  // diagnostics (ROBOVLN_GEMM_TIMES): [CTA][16] SM-clock stamps of the first tile's phases; null in production
  int dbg;           // timing experiments only (ROBOVLN_EPI_DEBUG bit mask; results are wrong when set)
};

struct GemmTcPlan {
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
  alignas(64) CUtensorMap tmC;
  alignas(64) CUtensorMap tmR;   // residual (plain GEMM, 32-row boxes); a copy of tmC when unused
  GemmTcParams p;
  int BN = 128;
  int ctas = 1;        // 2: CTA-pair (cta_group::2) kernel, 256 x BN tiles
  int grid = 0;
  bool valid = false;
  ConvGemm desc;       // kept for the SIMT validation path
};

// gemm_tc.cu
void gemm_tc_make_plan(const ConvGemm& g, GemmTcPlan* plan, int force_bn = 0);
void gemm_tc_launch(const GemmTcPlan& plan, cudaStream_t stream);
int  device_sm_count();
// gemm_simt.cu -- naive CUDA-core implementation of the same contract (validation only;
// selected with ROBOVLN_GEMM=simt).  Still a GPU kernel: there is no CPU path anywhere.
void gemm_simt_launch(const ConvGemm& g, cudaStream_t stream);
bool use_simt_gemm();

// elementwise.cu
void rgb_stem_im2col(const float* rgb, h16* out, int NB, int H, int W, int Kpitch, cudaStream_t s);
void rgb_pad_convert(const float* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s);
void rgb_pad_convert4(const float* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s);
void rgb_pad_convert4_u8(const uint8_t* rgb, h16* out, int NB, int H, int W, int Wp, cudaStream_t s);   // row-pair interleaved (packed stem)
void maxpool3x3s2(const h16* in, h16* out, int NB, int H, int W, int C, cudaStream_t s);
void depth_stem_conv(const float* depth, const float* w, h16* out, int NB, int H, int W, cudaStream_t s);
void gn_stats(const h16* x, float* stats, int NB, int HW, int C, int G, cudaStream_t s);
struct GnApply {
  const h16* x; const float* stats; const float* gamma; const float* beta;
  int NB, HW, C, G;
  int relu;
  int res_mode;                  // 0 none, 1 plain h16 residual, 2 group-normalised residual
  const h16* res; const float* res_stats; const float* res_gamma; const float* res_beta;
  h16* out; int64_t out_pitch;  // elements per pixel in the output (>= C)
};
void gn_apply(const GnApply& a, cudaStream_t s);
void gn_fused(const GnApply& a, cudaStream_t s);   // statistics + apply in one launch (a.stats / a.res_stats unused)
void rgb_pool(const h16* feat, int NB, int H, int W, int C, h16* tokens, int64_t tok_pitch, h16* cellmean,
              int64_t cm_pitch, h16* gmean, cudaStream_t s);
void fill_spatial_embedding(const float* emb_flat, h16* tokens, int NB, int64_t tok_pitch, int col0, h16* cellmean,
                            int64_t cm_pitch, cudaStream_t s);
void bert_embed_ln(const int64_t* ids_i64, const float* ids_f32, int id_rows, int R, int L, const float* word,
                   const float* pos, const float* type0, const float* g, const float* b, h16* out, cudaStream_t s);
void layernorm_rows(const float* x, int M, int D, const float* g, const float* b, float eps, const float* pe, int pe_rows,
                    h16* out, cudaStream_t s);
void token_mean(const h16* x, int n_mod, int B, int L, int D, h16* out, int64_t out_pitch, int64_t mod_stride,
                cudaStream_t s);
void sub_task_embed(const int64_t* ids, const float* table, int B, h16* out, int64_t out_pitch, cudaStream_t s);
void heads_linear(const float* y, int M, int K, const float* w, const float* b, int n_out, float* out, cudaStream_t s);
void heads_fused(const float* y, int M, int K, const float* wA, const float* bA, int nA, float* outA, const float* wB,
                 const float* bB, int nB, float* outB, int64_t* amax, const float* emb, h16* emb_out, int64_t emb_pitch,
                 cudaStream_t s);
void argmax_rows(const float* x, int M, int n, int64_t* out, cudaStream_t s);
void sinusoid_table(float* pe, int L, int D, cudaStream_t s);

// gemm_tc.cu: 2-D tensor map over a row-major 16-bit matrix [rows, cols] (pitch in bytes), 128B swizzle
void tma_encode_2d_h16(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_bytes, uint32_t box_cols,
                       uint32_t box_rows);
void tma_encode_nd_h16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box);
// vla_block.cu -- the fused cross-modal attention block (attention + fc_o + LN1 + FFN + LN2 + token mean), L <= 128
struct VlaBlock {
  int B = 0, L = 0, q_shared = 0;
  const h16* q0 = nullptr;        // [R*L, 256] = LN0(relu(ins_fc(bert))) + PE, R = 1 (q_shared) or B
  const h16* kvx = nullptr;       // [2*B*16, 1288]: per visual cell K'(4 heads x 256) | c(8) | V(256); rgb rows then depth rows
  int64_t kvx_pitch = 1288;
  const h16 *wo = nullptr, *w1 = nullptr, *w2 = nullptr;   // fc_o [256,256], fc1 [1024,256], fc2 [256,1024]
  const float *bo = nullptr, *b1 = nullptr, *b2 = nullptr, *ln1g = nullptr, *ln1b = nullptr, *ln2g = nullptr, *ln2b = nullptr;
  float eps = 1e-5f;
  h16* out = nullptr;             // pooled [B, out_pitch]; modality m at column m*256
  int64_t out_pitch = 512;
  h16* y_tokens = nullptr;        // parity tests only: token-level output [2, B, L, 256]
};
struct VlaBlockPlan {
  alignas(64) CUtensorMap tmQ0;
  alignas(64) CUtensorMap tmKp;
  alignas(64) CUtensorMap tmV;
  alignas(64) CUtensorMap tmWo;
  alignas(64) CUtensorMap tmW1;
  alignas(64) CUtensorMap tmW2;
  VlaBlock d;
  bool valid = false;
};
void vla_block_make_plan(const VlaBlock& d, VlaBlockPlan* plan);
void vla_block_launch(const VlaBlockPlan& plan, cudaStream_t s);
extern int g_vla_variant;   // 0: ROBOVLN_VLA_PAIR (default pair); 1: one CTA per (env, modality) tile; 2: CTA pair per environment
// attention_tc.cu -- tcgen05 / TMEM / TMA self-attention for L <= 128 (ROBOVLN_ATTN=tc selects it in the engine)
bool use_tc_attention();
void bert_self_attention_tc(const h16* qkv, h16* ctx, int R, int L, int heads, cudaStream_t s);
// attention.cu
void bert_self_attention(const h16* qkv, h16* ctx, int R, int L, int heads, cudaStream_t s);
void vla_cross_attention(const h16* q, const h16* kv, h16* ctx, int B, int L, int n_mod, int q_shared,
                         cudaStream_t s);

// prep.cu -- weight packing (BatchNorm fold + OIHW -> K-major h16), bulk buffer comparison, content checksum
void pack_weight(const float* w, const float* gamma, const float* beta, const float* mean, const float* var, float eps,
                 h16* out, float* bias_out, int O, int I, int KH, int KW, int64_t out_pitch, cudaStream_t s);
void compare_many(const void* const* a_dev, const void* const* b_dev, const long long* words_dev, int n, int* mismatch_dev,
                  cudaStream_t s);
void checksum(const void* p, size_t bytes, unsigned long long* out2_dev, cudaStream_t s);

// train.cu -- DAgger update tail: fused losses (value + gradient in one launch) and multi-tensor Adam / AdamW
void hi_loss(const float* logits, const float* oracle_f, const int64_t* oracle_i, int T, int C, float* loss_out, float* dlogits,
             cudaStream_t s);
void lo_loss(const float* actions, const float* corrected, const float* stop, const float* oracle_stop, int T, int A, float* loss_out,
             float* d_actions, float* d_stop, cudaStream_t s);
void fused_adam(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                const long long* numel, const long long* chunk_start, int n_tensors, long long total_chunks, double lr, double b1,
                double b2, double eps, double wd, int decoupled, double step_size, double bc2_sqrt, cudaStream_t s);
int adam_chunk_elems();

// lstm.cu
void lstm_forward(const float* gx, const h16* whh, const float* masks, int mask_stride, const float* hc_in,
                  float* hc_out, float* h_scratch, float* y, int T, int N, cudaStream_t s);

}  // namespace rvb
