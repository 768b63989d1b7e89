/*
 * robovln_b200 -- C ABI of the B200-native HCM policy forward pass.
 *
 * The reference (GT-RIPL/robo-vln) has no plugin / FFI layer: its seam for this path is the
 * Python nn.Module API consumed by robo_vln_baselines/hierarchical_trainer.py:50-51,506,539,
 * 1096-1100.  The Python classes in robo-vln_b200/ mirror that API and call the functions
 * below through ctypes; every entry point cites the reference code it replaces.
 *
 * Conventions
 *   - plain C types only; device pointers are raw addresses in the current CUDA context;
 *     `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no function synchronises the device or allocates device memory; all scratch lives in
 *     the caller-provided workspace given to hcm_plan();
 *   - return value 0 = success, otherwise a cudaError_t or -1; hcm_last_error() returns a
 *     thread-local message for the last failure;
 *   - rows of a batch are ordered t-major (row = t*N + n), exactly as
 *     RNNStateEncoder.seq_forward expects (habitat_baselines/rl/models/rnn_state_encoder.py:85-100).
 */
#ifndef ROBOVLN_B200_H_
#define ROBOVLN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hcm_engine hcm_engine;

enum { HCM_F32 = 0, HCM_BF16 = 1, HCM_I64 = 2, HCM_F16 = 3 };

/* The library is compiled for ONE 16-bit storage / tensor-core operand type:
 * librobovln_b200.so = IEEE fp16 (default), librobovln_b200_bf16.so = bfloat16.  Every
 * `*_h16` / 16-bit tensor argument below must have the type hcm_dtype() reports
 * (HCM_F16 or HCM_BF16); accumulation, statistics, softmax and the LSTM state are fp32. */
int hcm_dtype(void);

/* Shape of one forward call. */
typedef struct hcm_shape {
  int32_t B;          /* rows in the batch = T*N                                             */
  int32_t N;          /* environments (columns of the hidden state [2,N,512])                 */
  int32_t L;          /* instruction tokens                                                   */
  int32_t instr_rows; /* 1 (one instruction shared by all rows, expanded like
                         seq2seq_highlevel_cma.py:189-190) or B                               */
  int32_t rgb_h, rgb_w;     /* RGB frame size (reference default 224x224; BASELINE 256x256)   */
  int32_t depth_h, depth_w; /* depth frame size (256x256)                                     */
} hcm_shape;

const char* hcm_last_error(void);
const char* hcm_version(void);

/* ---- engine lifetime ------------------------------------------------------------------ */
int  hcm_create(hcm_engine** out);
void hcm_destroy(hcm_engine* e);

/* Register a prepared (kernel-layout) weight tensor living in device memory.  The engine keeps
 * the pointer, not a copy; the Python modules own the storage and re-register after
 * load_state_dict()/optimizer steps.  Names are listed in robo-vln_b200/weight_prep.py; they
 * derive from the reference state_dict keys. */
int hcm_set_tensor(hcm_engine* e, const char* name, const void* dev_ptr, int dtype, int ndim,
                   const int64_t* shape);
/* Declare which halves are present: the hi model (Seq2Seq_HighLevel_CMA), the lo model
 * (Seq2Seq_LowLevel), and whether lo's frozen trunks are bit-identical to hi's so that one
 * trunk pass serves both (SURVEY.md 7.2 "dedup legality"). */
int hcm_finalize_weights(hcm_engine* e, int have_hi, int have_lo, int lo_shares_trunks);

/* ---- planning -------------------------------------------------------------------------- */
size_t hcm_workspace_bytes(hcm_engine* e, const hcm_shape* shape);
int    hcm_plan(hcm_engine* e, const hcm_shape* shape, void* workspace, size_t workspace_bytes);

/* ---- forward --------------------------------------------------------------------------- */
/* Seq2Seq_HighLevel_CMA.forward (robo_vln_baselines/models/seq2seq_highlevel_cma.py:170-233).
 *   rgb [B,rgb_h,rgb_w,3] f32 0..255, depth [B,depth_h,depth_w,1] f32,
 *   instruction ids as f32 (what the trainer passes) OR i64 (exactly one non-NULL), [instr_rows,L],
 *   masks: element (row*mask_stride) is masks[row,0]; hc_in/hc_out [2,N,512] f32 (no aliasing),
 *   logits [B,4] f32. */
int hcm_forward_hi(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                   const int64_t* instr_i64, const float* masks, int mask_stride, const float* hc_in,
                   float* logits, float* hc_out, void* stream);

/* Seq2Seq_LowLevel.forward (robo_vln_baselines/models/seq2seq_lowlevel.py:116-162).
 *   sub_goal i64 [B] in 0..4; actions [B,2], stop_logit [B,1].
 *   reuse_trunks != 0: take the RGB/depth features computed by the preceding hcm_forward_hi on
 *   the same observations (only legal when hcm_finalize_weights was told lo_shares_trunks). */
int hcm_forward_lo(hcm_engine* e, const float* rgb, const float* depth, const float* masks,
                   int mask_stride, const int64_t* sub_goal, const float* hc_in, float* actions,
                   float* stop_logit, float* hc_out, int reuse_trunks, void* stream);

/* One rollout step of the hierarchy as hierarchical_trainer.py:1095-1101 runs it:
 * hi -> argmax over the 4 sub-goal logits -> lo, trunks evaluated once.
 *   sub_goal_out i64 [B] receives the argmax (may be NULL). */
int hcm_forward_policy(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                       const int64_t* instr_i64, const float* masks, int mask_stride,
                       const float* hc_hi_in, const float* hc_lo_in, float* logits, float* actions,
                       float* stop_logit, float* hc_hi_out, float* hc_lo_out, int64_t* sub_goal_out,
                       void* stream);

/* Same step with HOST buffers (pinned recommended): H2D of the observations, the forward, and
 * D2H of the outputs are all enqueued on `stream`; the call returns after synchronising it.
 * This is the end-to-end entry bench.py times. */
int hcm_forward_policy_host(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                            const float* masks, const float* hc_hi_in, const float* hc_lo_in,
                            float* logits, float* actions, float* stop_logit, float* hc_hi_out,
                            float* hc_lo_out, void* stream);

/* Element type behind the `rgb` pointer of every entry point: 0 = float32 in 0..255 (default; what the reference's
 * batch_obs produces, robo_vln_baselines/common/utils.py:59-118), 1 = uint8 [B,H,W,3] exactly as the RGB sensor
 * delivers it (habitat_extensions/config/robo_vln_task.yaml:10-13).  Results are identical; the uint8 form moves a
 * quarter of the bytes (observation ingest). */
int hcm_set_rgb_format(hcm_engine* e, int fmt);

/* Instruction cache (the reference re-tokenises and re-encodes the unchanged instruction at
 * every step, common/utils.py:87-118 + seq2seq_highlevel_cma.py:189-195).  skip = 1: the next forward calls keep the
 * BERT output and the query-side projection of the previous call (the caller guarantees that the instruction tokens
 * are identical and that a full forward has run since the last hcm_plan); skip = 0 (default): BERT runs. */
int hcm_set_skip_bert(hcm_engine* e, int skip);

/* Number of kernels the last forward call launched (for bench.py's gpu_launches). */
int64_t hcm_last_launch_count(hcm_engine* e);

/* Profiling replay of hcm_forward_policy on ONE stream with a CUDA event between consecutive
 * launches.  Writes a JSON array [{"name","ms","flops"}...] (flops = algorithmic 2*M*N*K of
 * tensor-core launches, 0 for the others) into json_out.  Synchronises `stream`. */
int hcm_profile_policy(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                       const int64_t* instr_i64, const float* masks, int mask_stride,
                       const float* hc_hi_in, const float* hc_lo_in, float* logits, float* actions,
                       float* stop_logit, float* hc_hi_out, float* hc_lo_out, char* json_out,
                       size_t json_cap, void* stream);

/* ---- stage entry points (parity tests, ncu) --------------------------------------------- */
/* Each runs one stage on the planned shape and leaves its result in an engine buffer that
 * hcm_get_buffer exposes (name -> device pointer, dtype, shape). */
int hcm_run_rgb_trunk(hcm_engine* e, const float* rgb, int use_lo_weights, void* stream);
int hcm_run_depth_trunk(hcm_engine* e, const float* depth, int use_lo_weights, void* stream);
int hcm_run_bert(hcm_engine* e, const float* instr_f32, const int64_t* instr_i64, void* stream);
/* The three frozen encoders of one call (RGB trunk, depth trunk and, if with_bert, BERT) on the
 * engine's stream fork; results in the buffers "rgb_tokens", "rgb_cellmean", "rgb_gmean",
 * "depth_tokens", "bert".  Used by the training path: the trainable tail then runs under
 * autograd on the host framework side (robo-vln_b200/torch_tail.py). */
int hcm_run_encoders(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                     const int64_t* instr_i64, int with_bert, int use_lo_weights, void* stream);
/* cross-modal block on caller tensors: bert [R*L,768] bf16 (R = B or 1 per the plan),
 * rgb_spatial / depth_spatial [B*16,256] bf16 (outputs of rgb_kv / depth_kv) ->
 * pooled [B, 512] bf16 (ins_rgb_att | ins_depth_att).  BASELINE.json configs[2]. */
int hcm_run_cross_modal(hcm_engine* e, const void* bert_bf16, const void* rgb_spatial_bf16,
                        const void* depth_spatial_bf16, void* pooled_bf16, void* stream);
int hcm_get_buffer(hcm_engine* e, const char* name, void** dev_ptr, int* dtype, int* ndim,
                   int64_t* shape /* [>=4] */);

/* Device-to-device copy of a stage buffer into caller memory (bytes must match exactly). */
int hcm_copy_buffer(hcm_engine* e, const char* name, void* dst_dev, size_t bytes, void* stream);

/* ---- kernel-level entry points (unit parity tests) -------------------------------------- */
/* Implicit-GEMM convolution / linear layer on NHWC bf16:
 *   out[m,n] = act(sum A(m,tap,c) W[n,tap*Cin+c] + bias[n] + res[m % res_rows, n]).
 * impl: 0 = tcgen05 kernel, 1 = CUDA-core validation kernel.  force_bn: 0 or 64/128/256. */
int rvb_conv_gemm(const void* in_bf16, int NB, int H, int W, int Cin, int64_t in_pitch,
                  const void* w_bf16, int Cout, int KH, int KW, int stride, int pad,
                  const float* bias, const void* res_bf16, int64_t ldr, int res_rows, int act,
                  void* out, int64_t ldc, int out_f32, int force_bn, int impl, int window,
                  int64_t win_row_pitch, void* stream);
/* window != 0 ("window mode", the RGB stem): `in` is the zero-padded [NB,H,win_row_pitch/8,8] image
 * written by rvb_rgb_pad_convert, W is the OUTPUT width, KW must be 1 and Cin 64 (8 px x 8 ch per
 * filter row), weights [Cout, KH*64]. */
/* GEMM with LayerNorm folded into the store: out[M,N] (h16) = LN(act(A[M,K] . W[N,K]^T + bias + res[m % res_rows]))
 * * gamma + beta (+ pe[m % pe_rows]); N in {256, 512, 768}; statistics in fp32 over the fp32 accumulators
 * (BertSelfOutput / BertOutput, modeling_bert; transformer.py:38-43,111-126,262-281). */
int rvb_gemm_ln(const void* a_h16, int64_t M, int K, const void* w_h16, int N, const float* bias, const void* res_h16,
                int res_rows, int act, const float* gamma, const float* beta, float eps, const float* pe, int pe_rows,
                void* out_h16, void* stream);
/* Convolution (or 1x1 GEMM) with GroupNorm folded into the store, for output maps of 16 or 64 pixels per sample:
 * out = relu?(GroupNorm_groups(conv(in, w)) * gamma + beta + res)  (habitat_baselines/rl/ddppo/policy/resnet.py:65-77,
 * layers 3-4 of the depth trunk).  Cout/groups in {8,16,32,64}. */
int rvb_conv_gemm_gn(const void* in_h16, int NB, int H, int W, int Cin, const void* w_h16, int Cout, int KH, int stride,
                     int pad, const float* gamma, const float* beta, int groups, int relu, const void* res_h16,
                     void* out_h16, void* stream);
int rvb_rgb_pad_convert(const float* rgb, void* out_h16, int NB, int H, int W, int Wp, void* stream);
/* Packed stem (window == 2, what the engine runs): zero-padded ROW-PAIR-INTERLEAVED image [NB, (H+6)/2, Wp, 2, 4]
 * (H even); the conv takes H = (H+6)/2 row pairs, W = OUTPUT width, in_pitch 8, KH = 4 (row pairs), KW 1, stride 2,
 * Cin 64, win_row_pitch = Wp*8, weights [Cout, 4*64] with k = (r/2)*64 + s*8 + (r%2)*4 + c. */
int rvb_rgb_pad_convert4(const float* rgb, void* out_h16, int NB, int H, int W, int Wp, void* stream);
int rvb_groupnorm(const void* x_bf16, float* stats /* [NB,G,2] zeroed by the call */, const float* gamma,
                  const float* beta, int NB, int HW, int C, int G, int relu, const void* res_bf16,
                  void* out_bf16, int64_t out_pitch, void* stream);
int rvb_layernorm(const float* x, int M, int D, const float* gamma, const float* beta, float eps,
                  const float* pe, int pe_rows, void* out_bf16, void* stream);
int rvb_bert_attention(const void* qkv_bf16, void* ctx_bf16, int R, int L, int heads, void* stream);
/* Same contract as rvb_bert_attention on the tcgen05 / TMEM / TMA kernel (L <= 128): S = Q K^T and O = P V as
 * tcgen05.mma with V consumed as an MN-major operand, softmax read straight out of TMEM. */
int rvb_bert_attention_tc(const void* qkv_h16, void* ctx_h16, int R, int L, int heads, void* stream);
int rvb_vla_attention(const void* q_bf16, const void* kv_bf16, void* ctx_bf16, int B, int L, int q_rows,
                      void* stream);
/* Fused cross-modal block (csrc/vla_block.cu; transformer.py:262-281, :209-221, :81-126, :38-43): attention over the 16
 * visual cells + fc_o + LayerNorm + position-wise FFN + LayerNorm + token mean in one tcgen05 kernel, L <= 128.
 *   q0  [R*L, 256]      LN0(relu(ins_fc(bert))) + PE, R = 1 (q_shared) or B
 *   kvx [2*B*16, 1288]  per visual cell (rgb rows, then depth rows): K'(4 heads x 256) | c(4 + 4 pad) | V(256), where
 *                       K'_h = Wq_h^T k_h and c_h = bq_h . k_h fold fc_q into the key side
 *   out [B, out_pitch]  pooled tokens, modality m at column m*256;  y_tokens (optional) [2, B, L, 256] */
int rvb_vla_block(const void* q0_h16, const void* kvx_h16, const void* wo_h16, const void* w1_h16, const void* w2_h16,
                  const float* bo, const float* b1, const float* b2, const float* ln1g, const float* ln1b,
                  const float* ln2g, const float* ln2b, float eps, int B, int L, int q_shared, void* out_h16,
                  int64_t out_pitch, void* y_tokens_h16, void* stream);
/* which implementation rvb_vla_block / the engine launch: 0 = default (the CTA-pair kernel; ROBOVLN_VLA_PAIR=0 selects the
 * other), 1 = one CTA per (environment, modality) tile, 2 = one 2-CTA cluster (cta_group::2 MMAs) per environment */
int rvb_vla_block_variant(int variant);
int rvb_lstm(const float* gx, const void* whh_bf16, const float* masks, int mask_stride,
             const float* hc_in, float* hc_out, float* h_scratch, float* y, int T, int N, void* stream);
int rvb_maxpool3x3s2(const void* in_bf16, void* out_bf16, int NB, int H, int W, int C, void* stream);
int rvb_rgb_stem_im2col(const float* rgb, void* out_bf16, int NB, int H, int W, int Kpitch, void* stream);
int rvb_depth_stem(const float* depth, const float* w, void* out_bf16, int NB, int H, int W, void* stream);

/* ---- DAgger update tail (csrc/train.cu; robo_vln_baselines/hierarchical_trainer.py:492-560) --------------------
 * rvb_hi_loss: logits.masked_fill_(oracle == 0, 0); CrossEntropyLoss(ignore_index=-1)(logits, oracle - 1) (:506-511).
 *   oracle = the vln_oracle_action_sensor column (float32 or int64, one of the two pointers); loss_out2 = {loss, #rows
 *   counted}; dlogits (optional) = d loss / d logits [T, C].
 * rvb_lo_loss: actions.masked_fill_(corrected == 0, 0); MSELoss + BCEWithLogitsLoss over oracle_stop != -1 (:539-553).
 *   loss_out3 = {action loss, stop loss, #stop rows counted}; d_actions [T, A] / d_stop [T] (optional) = gradient of their sum.
 * rvb_fused_adam: one launch updates n_tensors fp32 tensors (device arrays of device pointers; chunk_start = prefix sum
 *   of ceil(numel / rvb_adam_chunk_elems()), n_tensors + 1 entries).  decoupled = 1: torch.optim.AdamW (hi, :329);
 *   0: torch.optim.Adam with L2 weight decay (lo, :332).  step_size = lr / (1 - beta1^step),
 *   bias_correction2_sqrt = sqrt(1 - beta2^step), computed by the caller in double precision as torch does. */
int rvb_hi_loss(const float* logits, const float* oracle_f32, const int64_t* oracle_i64, int T, int C, float* loss_out2,
                float* dlogits, void* stream);
int rvb_lo_loss(const float* actions, const float* corrected, const float* stop_logit, const float* oracle_stop, int T, int A,
                float* loss_out3, float* d_actions, float* d_stop, void* stream);
int rvb_fused_adam(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const int64_t* numel, const int64_t* chunk_start, int n_tensors, int64_t total_chunks, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int decoupled, double step_size, double bias_correction2_sqrt,
                   void* stream);
int rvb_adam_chunk_elems(void);

/* ---- host-side plumbing kernels (csrc/prep.cu) --------------------------------------------
 * rvb_pack_weight: fp32 parameter [O, I, KH, KW] (nn.Conv2d / nn.Linear layout, the reference's state_dict) ->
 *   16-bit [O, KH*KW*I] with k = (r*KW + s)*I + c, rows `out_pitch` elements apart (0 = dense); with bn_* given,
 *   eval-mode BatchNorm (torchvision ResNet-50 behind resnet_encoders.py:151) is folded in: weights scaled by
 *   gamma / sqrt(var + eps), bias_out[o] = beta - mean * scale.
 * rvb_compare_many: mismatch_dev[0] = 1 iff any of the n buffer pairs (device arrays of device pointers, sizes in
 *   32-bit words) differ -- the bit-identity check that makes running hi's and lo's frozen trunks once legal.
 * rvb_checksum: out2_dev[0..1] = 128-bit order-independent content checksum of a device buffer (used to decide
 *   whether lo may reuse the trunk features hi computed: hierarchical_trainer.py:1096-1100 hands both models
 *   the same observation batch). */
int rvb_pack_weight(const float* w_f32, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                    const float* bn_var, float eps, void* out_h16, float* bias_out, int O, int I, int KH, int KW,
                    int64_t out_pitch, void* stream);
int rvb_compare_many(const void* const* a_dev, const void* const* b_dev, const int64_t* words_dev, int n,
                     int* mismatch_dev, void* stream);
int rvb_checksum(const void* dev_ptr, size_t bytes, uint64_t* out2_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ROBOVLN_B200_H_ */
