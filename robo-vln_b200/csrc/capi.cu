// extern "C" boundary (include/robovln_b200.h): exception -> error code translation only.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"

using namespace rvb;

struct hcm_engine {
  Engine eng;
};

namespace {
thread_local std::string g_last_error;

template <class F>
int guarded(F&& f) {
  try {
    g_last_error.clear();
    f();
    return 0;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code != 0 ? e.code : -1;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return -1;
  }
}
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline const h16* B16(const void* p) { return reinterpret_cast<const h16*>(p); }
inline h16* B16(void* p) { return reinterpret_cast<h16*>(p); }
}  // namespace

extern "C" {

const char* hcm_last_error(void) { return g_last_error.c_str(); }
const char* hcm_version(void) { return "robovln_b200 0.1 (sm_100a: tcgen05/TMEM/TMA; 16-bit type " RVB_H16_NAME ")"; }
int hcm_dtype(void) { return RVB_H16_CODE; }

int hcm_create(hcm_engine** out) {
  return guarded([&] {
    RVB_CHECK(out != nullptr, "hcm_create: null out");
    int dev = 0;
    RVB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RVB_CUDA(cudaGetDeviceProperties(&prop, dev));
    RVB_CHECK(prop.major == 10, std::string("robovln_b200 needs an sm_100 GPU (B200); found sm_") +
                                    std::to_string(prop.major) + std::to_string(prop.minor));
    hcm_engine* e = new hcm_engine();
    const char* ms = std::getenv("ROBOVLN_MULTISTREAM");
    e->eng.multi_stream_ = (ms == nullptr) ? true : (std::strcmp(ms, "0") != 0);
    *out = e;
  });
}

void hcm_destroy(hcm_engine* e) { delete e; }

int hcm_set_tensor(hcm_engine* e, const char* name, const void* dev_ptr, int dtype, int ndim, const int64_t* shape) {
  return guarded([&] { e->eng.set_tensor(name, dev_ptr, dtype, ndim, shape); });
}

int hcm_finalize_weights(hcm_engine* e, int have_hi, int have_lo, int lo_shares_trunks) {
  return guarded([&] { e->eng.finalize(have_hi, have_lo, lo_shares_trunks); });
}

size_t hcm_workspace_bytes(hcm_engine* e, const hcm_shape* shape) {
  size_t n = 0;
  int rc = guarded([&] { n = e->eng.plan(*shape, nullptr, 0); });
  return rc == 0 ? n : 0;
}

int hcm_plan(hcm_engine* e, const hcm_shape* shape, void* workspace, size_t workspace_bytes) {
  return guarded([&] {
    RVB_CHECK(workspace != nullptr, "hcm_plan: null workspace");
    // deterministic contents for padding columns that no kernel writes
    RVB_CUDA(cudaMemset(workspace, 0, workspace_bytes));
    e->eng.plan(*shape, workspace, workspace_bytes);
    // the memsets above ran on the legacy default stream; the first forward may be issued on a non-blocking
    // stream that is not ordered after it
    RVB_CUDA(cudaDeviceSynchronize());
  });
}

int hcm_forward_hi(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                   const int64_t* instr_i64, const float* masks, int mask_stride, const float* hc_in, float* logits,
                   float* hc_out, void* stream) {
  return guarded([&] {
    RunArgs a;
    a.rgb = rgb; a.depth = depth; a.instr_f32 = instr_f32; a.instr_i64 = instr_i64;
    a.masks = masks; a.mask_stride = mask_stride; a.hc_hi_in = hc_in; a.hc_hi_out = hc_out; a.logits = logits;
    e->eng.args_ = a;
    e->eng.forward_hi_graphed(S(stream));
  });
}

int hcm_forward_lo(hcm_engine* e, const float* rgb, const float* depth, const float* masks, int mask_stride,
                   const int64_t* sub_goal, const float* hc_in, float* actions, float* stop_logit, float* hc_out,
                   int reuse_trunks, void* stream) {
  return guarded([&] {
    RunArgs a;
    a.rgb = rgb; a.depth = depth; a.masks = masks; a.mask_stride = mask_stride; a.sub_goal = sub_goal;
    a.hc_lo_in = hc_in; a.hc_lo_out = hc_out; a.actions = actions; a.stop = stop_logit;
    e->eng.args_ = a;
    e->eng.forward_lo(reuse_trunks != 0, S(stream));
  });
}

int hcm_forward_policy(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                       const int64_t* instr_i64, const float* masks, int mask_stride, const float* hc_hi_in,
                       const float* hc_lo_in, float* logits, float* actions, float* stop_logit, float* hc_hi_out,
                       float* hc_lo_out, int64_t* sub_goal_out, void* stream) {
  return guarded([&] {
    RunArgs a;
    a.rgb = rgb; a.depth = depth; a.instr_f32 = instr_f32; a.instr_i64 = instr_i64;
    a.masks = masks; a.mask_stride = mask_stride;
    a.hc_hi_in = hc_hi_in; a.hc_lo_in = hc_lo_in; a.hc_hi_out = hc_hi_out; a.hc_lo_out = hc_lo_out;
    a.logits = logits; a.actions = actions; a.stop = stop_logit; a.sub_goal_out = sub_goal_out;
    e->eng.args_ = a;
    e->eng.forward_policy_graphed(S(stream));
  });
}

int hcm_forward_policy_host(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                            const float* masks, const float* hc_hi_in, const float* hc_lo_in, float* logits,
                            float* actions, float* stop_logit, float* hc_hi_out, float* hc_lo_out, void* stream) {
  return guarded([&] {
    e->eng.forward_policy_host(rgb, depth, instr_f32, masks, hc_hi_in, hc_lo_in, logits, actions, stop_logit, hc_hi_out,
                               hc_lo_out, S(stream));
  });
}

int hcm_set_rgb_format(hcm_engine* e, int fmt) {
  return guarded([&] {
    RVB_CHECK(fmt == 0 || fmt == 1, "hcm_set_rgb_format: 0 = float32 (0..255), 1 = uint8");
    e->eng.rgb_fmt_ = fmt;
  });
}

int hcm_set_skip_bert(hcm_engine* e, int skip) {
  return guarded([&] { e->eng.skip_bert_ = skip != 0; });
}

int64_t hcm_last_launch_count(hcm_engine* e) { return e->eng.launches_; }

int hcm_profile_policy(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                       const int64_t* instr_i64, const float* masks, int mask_stride, const float* hc_hi_in,
                       const float* hc_lo_in, float* logits, float* actions, float* stop_logit, float* hc_hi_out,
                       float* hc_lo_out, char* json_out, size_t json_cap, void* stream) {
  return guarded([&] {
    RunArgs a;
    a.rgb = rgb; a.depth = depth; a.instr_f32 = instr_f32; a.instr_i64 = instr_i64;
    a.masks = masks; a.mask_stride = mask_stride;
    a.hc_hi_in = hc_hi_in; a.hc_lo_in = hc_lo_in; a.hc_hi_out = hc_hi_out; a.hc_lo_out = hc_lo_out;
    a.logits = logits; a.actions = actions; a.stop = stop_logit;
    e->eng.args_ = a;
    std::vector<OpTiming> t = e->eng.profile_policy(S(stream));
    std::string js = "[";
    for (size_t i = 0; i < t.size(); ++i) {
      char buf[512];
      snprintf(buf, sizeof(buf), "%s{\"name\":\"%s\",\"ms\":%.6f,\"flops\":%.1f,\"ctas\":%d}", i ? "," : "", t[i].name.c_str(),
               t[i].ms, t[i].flops, t[i].ctas);
      js += buf;
    }
    js += "]";
    RVB_CHECK(js.size() + 1 <= json_cap, "hcm_profile_policy: output buffer too small");
    std::memcpy(json_out, js.c_str(), js.size() + 1);
  });
}

int hcm_run_rgb_trunk(hcm_engine* e, const float* rgb, int use_lo_weights, void* stream) {
  return guarded([&] {
    RVB_CHECK(e->eng.planned_, "engine not planned");
    e->eng.args_.rgb = rgb;
    Stage& st = (use_lo_weights && !e->eng.st_rgb_lo_.empty()) ? e->eng.st_rgb_lo_ : e->eng.st_rgb_;
    e->eng.launches_ = e->eng.run(e->eng.st_pre_, S(stream));
    e->eng.launches_ += e->eng.run(st, S(stream));
  });
}

int hcm_run_depth_trunk(hcm_engine* e, const float* depth, int use_lo_weights, void* stream) {
  return guarded([&] {
    RVB_CHECK(e->eng.planned_, "engine not planned");
    e->eng.args_.depth = depth;
    Stage& st = (use_lo_weights && !e->eng.st_depth_lo_.empty()) ? e->eng.st_depth_lo_ : e->eng.st_depth_;
    e->eng.launches_ = e->eng.run(e->eng.st_pre_, S(stream));
    e->eng.launches_ += e->eng.run(st, S(stream));
  });
}

int hcm_run_encoders(hcm_engine* e, const float* rgb, const float* depth, const float* instr_f32,
                     const int64_t* instr_i64, int with_bert, int use_lo_weights, void* stream) {
  return guarded([&] {
    RVB_CHECK(e->eng.planned_, "engine not planned");
    RVB_CHECK(rgb != nullptr && depth != nullptr, "hcm_run_encoders: null observation");
    RVB_CHECK(!with_bert || (e->eng.have_hi_ && (instr_f32 != nullptr || instr_i64 != nullptr)),
              "hcm_run_encoders: BERT needs the hi weights and an instruction");
    e->eng.args_.rgb = rgb;
    e->eng.args_.depth = depth;
    e->eng.args_.instr_f32 = instr_f32;
    e->eng.args_.instr_i64 = instr_i64;
    e->eng.launches_ = e->eng.run(e->eng.st_pre_, S(stream));
    e->eng.run_encoders(with_bert != 0, use_lo_weights != 0, S(stream));
  });
}

int hcm_run_bert(hcm_engine* e, const float* instr_f32, const int64_t* instr_i64, void* stream) {
  return guarded([&] {
    RVB_CHECK(e->eng.planned_ && e->eng.have_hi_, "engine not planned");
    e->eng.args_.instr_f32 = instr_f32;
    e->eng.args_.instr_i64 = instr_i64;
    e->eng.launches_ = e->eng.run(e->eng.st_bert_, S(stream));
  });
}

int hcm_run_cross_modal(hcm_engine* e, const void* bert_bf16, const void* rgb_spatial_bf16,
                        const void* depth_spatial_bf16, void* pooled_bf16, void* stream) {
  return guarded([&] { e->eng.run_cross_modal(bert_bf16, rgb_spatial_bf16, depth_spatial_bf16, pooled_bf16, S(stream)); });
}

int hcm_get_buffer(hcm_engine* e, const char* name, void** dev_ptr, int* dtype, int* ndim, int64_t* shape) {
  return guarded([&] {
    std::vector<int64_t> shp;
    RVB_CHECK(e->eng.get_buffer(name, dev_ptr, dtype, &shp), std::string("unknown buffer '") + name + "'");
    *ndim = static_cast<int>(shp.size());
    for (size_t i = 0; i < shp.size(); ++i) shape[i] = shp[i];
  });
}

int hcm_copy_buffer(hcm_engine* e, const char* name, void* dst_dev, size_t bytes, void* stream) {
  return guarded([&] {
    std::vector<int64_t> shp;
    void* src = nullptr;
    int dtype = 0;
    RVB_CHECK(e->eng.get_buffer(name, &src, &dtype, &shp), std::string("unknown buffer '") + name + "'");
    size_t n = dtype == HCM_F32 ? 4 : (dtype == HCM_I64 ? 8 : 2);
    for (auto v : shp) n *= static_cast<size_t>(v);
    RVB_CHECK(bytes == n, "hcm_copy_buffer: size mismatch");
    RVB_CUDA(cudaMemcpyAsync(dst_dev, src, n, cudaMemcpyDeviceToDevice, S(stream)));
  });
}

// ---- kernel-level entry points -------------------------------------------------------------
int rvb_conv_gemm(const void* in_bf16, int NB, int H, int W, int Cin, int64_t in_pitch, const void* w_bf16, int Cout,
                  int KH, int KW, int stride, int pad, const float* bias, const void* res_bf16, int64_t ldr,
                  int res_rows, int act, void* out, int64_t ldc, int out_f32, int force_bn, int impl, int window,
                  int64_t win_row_pitch, void* stream) {
  return guarded([&] {
    ConvGemm g;
    g.window = window; g.win_row_pitch = win_row_pitch;
    g.in = B16(in_bf16); g.NB = NB; g.H = H; g.W = W; g.Cin = Cin; g.in_pitch = in_pitch;
    g.w = B16(w_bf16); g.Cout = Cout; g.KH = KH; g.KW = KW; g.stride = stride; g.pad = pad;
    g.bias = bias; g.res = B16(res_bf16); g.ldr = ldr; g.res_rows = res_rows; g.act = act;
    g.out = out; g.ldc = ldc; g.out_f32 = out_f32;
    if (impl == 1) {
      gemm_simt_launch(g, S(stream));
    } else {
      GemmTcPlan plan;
      gemm_tc_make_plan(g, &plan, force_bn);
      gemm_tc_launch(plan, S(stream));
    }
  });
}

int rvb_gemm_ln(const void* a_h16, int64_t M, int K, const void* w_h16, int N, const float* bias, const void* res_h16,
                int res_rows, int act, const float* gamma, const float* beta, float eps, const float* pe, int pe_rows,
                void* out_h16, void* stream) {
  return guarded([&] {
    ConvGemm g;
    g.in = B16(a_h16); g.NB = 1; g.H = 1; g.W = static_cast<int>(M); g.Cin = K; g.in_pitch = K;
    g.w = B16(w_h16); g.Cout = N; g.KH = g.KW = 1; g.stride = 1; g.pad = 0;
    g.bias = bias; g.res = B16(res_h16); g.ldr = N; g.res_rows = res_rows; g.act = act;
    g.out = out_h16; g.ldc = N; g.out_f32 = 0;
    g.ln_gamma = gamma; g.ln_beta = beta; g.ln_eps = eps; g.ln_pe = pe; g.ln_pe_rows = pe_rows > 0 ? pe_rows : 1;
    // ROBOVLN_LN_XCHG=global: exchange the row statistics of N > 256 through global memory instead of a cluster
    static const char* xe = std::getenv("ROBOVLN_LN_XCHG");
    if (xe != nullptr && std::strcmp(xe, "global") == 0 && N > 256) {
      static uint8_t* scratch = nullptr;
      const size_t m_tiles = static_cast<size_t>((M + 127) / 128);
      const size_t ws_bytes = m_tiles * 3 * 256 * 8, cap = size_t(64) << 20;
      RVB_CHECK(ws_bytes + m_tiles * 16 + 256 <= cap, "rvb_gemm_ln: M too large for the test scratch");
      if (scratch == nullptr) {
        RVB_CUDA(cudaMalloc(&scratch, cap));
        RVB_CUDA(cudaMemset(scratch, 0, cap));
      }
      g.ln_ws = scratch + (size_t(1) << 20);        // counters live in the first MiB
      g.ln_cnt = reinterpret_cast<int*>(scratch);
      // the engine gives every plan its own (monotonic, never reset) counters; this shared test scratch serves
      // calls with different N / M, so it is re-zeroed in stream order before every launch
      RVB_CUDA(cudaMemsetAsync(scratch, 0, m_tiles * 16, S(stream)));
    }
    GemmTcPlan plan;
    gemm_tc_make_plan(g, &plan, 0);
    gemm_tc_launch(plan, S(stream));
  });
}

int rvb_conv_gemm_gn(const void* in_h16, int NB, int H, int W, int Cin, const void* w_h16, int Cout, int KH, int stride,
                     int pad, const float* gamma, const float* beta, int groups, int relu, const void* res_h16,
                     void* out_h16, void* stream) {
  return guarded([&] {
    ConvGemm g;
    g.in = B16(in_h16); g.NB = NB; g.H = H; g.W = W; g.Cin = Cin; g.in_pitch = Cin;
    g.w = B16(w_h16); g.Cout = Cout; g.KH = g.KW = KH; g.stride = stride; g.pad = pad;
    g.res = B16(res_h16); g.ldr = Cout; g.res_rows = 0; g.act = relu ? ACT_RELU : ACT_NONE;
    g.out = out_h16; g.ldc = Cout; g.out_f32 = 0;
    g.gn_gamma = gamma; g.gn_beta = beta; g.gn_groups = groups; g.gn_hw = g.Ho() * g.Wo();
    GemmTcPlan plan;
    gemm_tc_make_plan(g, &plan, 0);
    gemm_tc_launch(plan, S(stream));
  });
}

int rvb_groupnorm(const void* x_bf16, float* stats, const float* gamma, const float* beta, int NB, int HW, int C, int G,
                  int relu, const void* res_bf16, void* out_bf16, int64_t out_pitch, void* stream) {
  return guarded([&] {
    if (stats == nullptr) {   // fused statistics + apply (what the engine runs)
      GnApply a{B16(x_bf16), nullptr, gamma, beta, NB, HW, C, G, relu, res_bf16 != nullptr ? 1 : 0, B16(res_bf16),
                nullptr, nullptr, nullptr, B16(out_bf16), out_pitch};
      gn_fused(a, S(stream));
      return;
    }
    RVB_CUDA(cudaMemsetAsync(stats, 0, static_cast<size_t>(NB) * G * 2 * sizeof(float), S(stream)));
    gn_stats(B16(x_bf16), stats, NB, HW, C, G, S(stream));
    GnApply a{B16(x_bf16), stats, gamma, beta, NB, HW, C, G, relu, res_bf16 != nullptr ? 1 : 0, B16(res_bf16),
              nullptr, nullptr, nullptr, B16(out_bf16), out_pitch};
    gn_apply(a, S(stream));
  });
}

int rvb_layernorm(const float* x, int M, int D, const float* gamma, const float* beta, float eps, const float* pe,
                  int pe_rows, void* out_bf16, void* stream) {
  return guarded([&] { layernorm_rows(x, M, D, gamma, beta, eps, pe, pe_rows > 0 ? pe_rows : 1, B16(out_bf16), S(stream)); });
}

int rvb_bert_attention(const void* qkv_bf16, void* ctx_bf16, int R, int L, int heads, void* stream) {
  return guarded([&] { bert_self_attention(B16(qkv_bf16), B16(ctx_bf16), R, L, heads, S(stream)); });
}

int rvb_bert_attention_tc(const void* qkv_h16, void* ctx_h16, int R, int L, int heads, void* stream) {
  return guarded([&] { bert_self_attention_tc(B16(qkv_h16), B16(ctx_h16), R, L, heads, S(stream)); });
}

int rvb_vla_attention(const void* q_bf16, const void* kv_bf16, void* ctx_bf16, int B, int L, int q_rows, void* stream) {
  return guarded([&] { vla_cross_attention(B16(q_bf16), B16(kv_bf16), B16(ctx_bf16), B, L, 1, q_rows == L ? 1 : 0, S(stream)); });
}

int rvb_vla_block(const void* q0_h16, const void* kvx_h16, const void* wo_h16, const void* w1_h16, const void* w2_h16,
                  const float* bo, const float* b1, const float* b2, const float* ln1g, const float* ln1b, const float* ln2g,
                  const float* ln2b, float eps, int B, int L, int q_shared, void* out_h16, int64_t out_pitch,
                  void* y_tokens_h16, void* stream) {
  return guarded([&] {
    VlaBlock d;
    d.B = B; d.L = L; d.q_shared = q_shared;
    d.q0 = B16(q0_h16); d.kvx = B16(kvx_h16); d.kvx_pitch = 1288;
    d.wo = B16(wo_h16); d.w1 = B16(w1_h16); d.w2 = B16(w2_h16);
    d.bo = bo; d.b1 = b1; d.b2 = b2; d.ln1g = ln1g; d.ln1b = ln1b; d.ln2g = ln2g; d.ln2b = ln2b;
    d.eps = eps; d.out = B16(out_h16); d.out_pitch = out_pitch; d.y_tokens = B16(y_tokens_h16);
    VlaBlockPlan plan;
    vla_block_make_plan(d, &plan);
    vla_block_launch(plan, S(stream));
  });
}

int rvb_vla_block_variant(int variant) {
  return guarded([&] {
    RVB_CHECK(variant >= 0 && variant <= 2, "rvb_vla_block_variant: 0 = default, 1 = single-CTA tiles, 2 = CTA pair");
    g_vla_variant = variant;
  });
}

int rvb_hi_loss(const float* logits, const float* oracle_f32, const int64_t* oracle_i64, int T, int C, float* loss_out2,
                float* dlogits, void* stream) {
  return guarded([&] { hi_loss(logits, oracle_f32, oracle_i64, T, C, loss_out2, dlogits, S(stream)); });
}

int rvb_lo_loss(const float* actions, const float* corrected, const float* stop_logit, const float* oracle_stop, int T, int A,
                float* loss_out3, float* d_actions, float* d_stop, void* stream) {
  return guarded([&] { lo_loss(actions, corrected, stop_logit, oracle_stop, T, A, loss_out3, d_actions, d_stop, S(stream)); });
}

int rvb_fused_adam(void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                   const int64_t* numel, const int64_t* chunk_start, int n_tensors, int64_t total_chunks, double lr, double beta1,
                   double beta2, double eps, double weight_decay, int decoupled, double step_size, double bias_correction2_sqrt,
                   void* stream) {
  return guarded([&] {
    fused_adam(reinterpret_cast<float* const*>(params), reinterpret_cast<const float* const*>(grads),
               reinterpret_cast<float* const*>(exp_avg), reinterpret_cast<float* const*>(exp_avg_sq),
               reinterpret_cast<const long long*>(numel), reinterpret_cast<const long long*>(chunk_start), n_tensors, total_chunks, lr,
               beta1, beta2, eps, weight_decay, decoupled, step_size, bias_correction2_sqrt, S(stream));
  });
}

int rvb_adam_chunk_elems(void) { return adam_chunk_elems(); }

int rvb_lstm(const float* gx, const void* whh_bf16, const float* masks, int mask_stride, const float* hc_in,
             float* hc_out, float* h_scratch, float* y, int T, int N, void* stream) {
  return guarded([&] { lstm_forward(gx, B16(whh_bf16), masks, mask_stride, hc_in, hc_out, h_scratch, y, T, N, S(stream)); });
}

int rvb_maxpool3x3s2(const void* in_bf16, void* out_bf16, int NB, int H, int W, int C, void* stream) {
  return guarded([&] { maxpool3x3s2(B16(in_bf16), B16(out_bf16), NB, H, W, C, S(stream)); });
}

int rvb_rgb_stem_im2col(const float* rgb, void* out_bf16, int NB, int H, int W, int Kpitch, void* stream) {
  return guarded([&] { rgb_stem_im2col(rgb, B16(out_bf16), NB, H, W, Kpitch, S(stream)); });
}

int rvb_rgb_pad_convert(const float* rgb, void* out_h16, int NB, int H, int W, int Wp, void* stream) {
  return guarded([&] { rgb_pad_convert(rgb, B16(out_h16), NB, H, W, Wp, S(stream)); });
}

int rvb_rgb_pad_convert4(const float* rgb, void* out_h16, int NB, int H, int W, int Wp, void* stream) {
  return guarded([&] { rgb_pad_convert4(rgb, B16(out_h16), NB, H, W, Wp, S(stream)); });
}

int rvb_pack_weight(const float* w_f32, const float* bn_gamma, const float* bn_beta, const float* bn_mean,
                    const float* bn_var, float eps, void* out_h16, float* bias_out, int O, int I, int KH, int KW,
                    int64_t out_pitch, void* stream) {
  return guarded([&] { pack_weight(w_f32, bn_gamma, bn_beta, bn_mean, bn_var, eps, B16(out_h16), bias_out, O, I, KH, KW, out_pitch, S(stream)); });
}

int rvb_compare_many(const void* const* a_dev, const void* const* b_dev, const int64_t* words_dev, int n, int* mismatch_dev,
                     void* stream) {
  return guarded([&] { compare_many(a_dev, b_dev, reinterpret_cast<const long long*>(words_dev), n, mismatch_dev, S(stream)); });
}

int rvb_checksum(const void* dev_ptr, size_t bytes, uint64_t* out2_dev, void* stream) {
  return guarded([&] { checksum(dev_ptr, bytes, reinterpret_cast<unsigned long long*>(out2_dev), S(stream)); });
}

int rvb_depth_stem(const float* depth, const float* w, void* out_bf16, int NB, int H, int W, void* stream) {
  return guarded([&] { depth_stem_conv(depth, w, B16(out_bf16), NB, H, W, S(stream)); });
}

}  // extern "C"
