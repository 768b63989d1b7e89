// Engine class behind the opaque hcm_engine handle of include/robovln_b200.h.
#pragma once

#include <functional>
#include <initializer_list>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/robovln_b200.h"
#include "rvb.h"

namespace rvb {

struct WTensor {
  const void* ptr = nullptr;
  int dtype = 0;  // HCM_F32 / HCM_BF16 / HCM_I64
  std::vector<int64_t> shape;
};

// Pointers that change from call to call; the planned closures read them at launch time.
struct RunArgs {
  const float* rgb = nullptr;
  const float* depth = nullptr;
  const float* instr_f32 = nullptr;
  const int64_t* instr_i64 = nullptr;
  const float* masks = nullptr;
  int mask_stride = 2;
  const int64_t* sub_goal = nullptr;
  const float* hc_hi_in = nullptr;
  const float* hc_lo_in = nullptr;
  float* hc_hi_out = nullptr;
  float* hc_lo_out = nullptr;
  float* logits = nullptr;
  float* actions = nullptr;
  float* stop = nullptr;
  int64_t* sub_goal_out = nullptr;
};

// One planned launch (or a short fixed sequence of launches); fn returns the kernel count.
struct Op {
  std::function<int(cudaStream_t)> fn;
  std::string name;     // filled for the profile listing
  double flops = 0.0;   // algorithmic FLOPs (2*M*N*K) for tensor-core ops, 0 otherwise
  int ctas = 0;         // persistent grid of a tensor-core op (CTAs = SMs it occupies), 0 when not recorded
  Op() = default;
  template <class F, class = typename std::enable_if<!std::is_same<typename std::decay<F>::type, Op>::value>::type>
  Op(F&& f) : fn(std::forward<F>(f)) {}
  int operator()(cudaStream_t s) const { return fn(s); }
};
using Stage = std::vector<Op>;

struct OpTiming {
  std::string name;
  double ms = 0.0;
  double flops = 0.0;
  int ctas = 0;
};

class Engine {
 public:
  Engine() = default;
  ~Engine();

  void set_tensor(const std::string& name, const void* ptr, int dtype, int ndim, const int64_t* shape);
  void finalize(int have_hi, int have_lo, int lo_shares_trunks);
  // workspace == nullptr: dry run, returns the bytes needed
  size_t plan(const hcm_shape& shp, void* workspace, size_t bytes);

  void forward_hi(cudaStream_t s);
  void forward_lo(bool reuse_trunks, cudaStream_t s);
  void forward_policy(cudaStream_t s);
  // forward_policy replayed from a CUDA graph (captured once per distinct set of input pointers; outputs go
  // through engine-owned buffers and are copied to the caller's).  ROBOVLN_GRAPH=0 disables.
  void forward_policy_graphed(cudaStream_t s);
  void forward_hi_graphed(cudaStream_t s);     // the module API's hi call (hierarchical_trainer.py:1096-1097), same cache
  void forward_graphed(int kind, cudaStream_t s);
  void forward_policy_host(const float* rgb, const float* depth, const float* instr, const float* masks,
                           const float* hc_hi_in, const float* hc_lo_in, float* logits, float* actions, float* stop,
                           float* hc_hi_out, float* hc_lo_out, cudaStream_t s);
  void run_cross_modal(const void* bert, const void* rgb_sp, const void* depth_sp, void* pooled, cudaStream_t s);
  // posts_hi / posts_lo: also run the tail ops that depend on ONE encoder only (rgb_kv, rgb_linear,
  // depth_kv, depth_linear, the query side of the cross-modal block; lo's rgb_fc / depth_fc) on
  // that encoder's stream, off the serial tail
  void run_encoders(bool with_bert, bool lo_weights, cudaStream_t s, bool posts_hi = false, bool posts_lo = false);
  // single-stream replay of forward_policy with a CUDA event between every op
  std::vector<OpTiming> profile_policy(cudaStream_t s);
  int run(const Stage& st, cudaStream_t s);
  bool get_buffer(const std::string& name, void** ptr, int* dtype, std::vector<int64_t>* shape) const;

  // 0: RGB frames are float32 in 0..255 (the reference's batch_obs output); 1: uint8 (the sensor's own format --
  // a quarter of the upload bytes).  Applies to the `rgb` pointer of every entry point.
  int rgb_fmt_ = 0;
  // Instruction cache: the caller asserts that the instruction tokens are the ones of the
  // previous call, so BERT (54 % of the step's FLOPs) and the query-side projection keep their outputs from that call.
  int extra_rounds_ = 0;   // persistent-grid policy of the stage being planned (ConvGemm::extra_rounds)
  int grid_cap_ = 0;   // > 0 while a stage whose GEMM grids are capped is being planned
  bool skip_bert_ = false;
  RunArgs args_;
  // forward_policy only: the hi head also writes argmax -> sub-goal ids and lo's sub-task embedding
  int64_t* policy_sg_ = nullptr;
  int64_t launches_ = 0;
  bool multi_stream_ = false;
  std::function<void(cudaStream_t)> before_rgb_;   // enqueued right before the RGB trunk (host-entry upload)
  bool planned_ = false;
  bool have_hi_ = false, have_lo_ = false, lo_shares_trunks_ = false;
  hcm_shape shp_{};
  Stage st_rgb_, st_depth_, st_rgb_lo_, st_depth_lo_, st_bert_, st_pre_, st_hi_tail_, st_lo_tail_, st_cm_only_;
  // single-encoder consumers, issued on the encoder's own stream before the join
  Stage st_rgb_post_hi_, st_depth_post_hi_, st_bert_post_, st_rgb_post_lo_, st_depth_post_lo_;

 private:
  const WTensor& W(const std::string& name, int dtype, std::initializer_list<int64_t> shape) const;
  const h16* Wb(const std::string& n, std::initializer_list<int64_t> s) const;
  const float* Wf(const std::string& n, std::initializer_list<int64_t> s) const;
  void* alloc(size_t bytes);
  float* new_stats(int G);
  GemmTcPlan* add_gemm(Stage& st, const ConvGemm& g, int force_bn = 0);
  static void label(Stage& st, const std::string& prefix);
  static ConvGemm linear(const h16* in, int64_t M, int K, int64_t lda, const h16* w, int N, const float* bias, int act,
                         void* out, int64_t ldc, int out_f32, const h16* res = nullptr, int64_t ldr = 0,
                         int res_rows = 0);
  void plan_rgb_trunk(const std::string& ns, Stage& st);
  h16* plan_rgb_blocks(const std::string& ns, Stage& st, h16* x, int NB, int& h, int& w, int& cin, int li_begin, int li_end,
                       h16* out_last);
  void plan_depth_trunk(const std::string& ns, Stage& st);
  void plan_bert(Stage& st);
  void plan_cross_modal(Stage& stq, Stage& st, const h16* bert, const h16* kvin, h16* out, int64_t out_pitch,
                        Stage* st_vis_rgb = nullptr, Stage* st_vis_depth = nullptr);
  void plan_hi_tail(Stage& pre, Stage& st);
  void plan_lo_tail(Stage& st);

  std::unordered_map<std::string, WTensor> weights_;
  bool finalized_ = false;
  bool dry_ = true;
  uint8_t* arena_base_ = nullptr;
  size_t arena_off_ = 0, arena_cap_ = 0;
  std::vector<std::unique_ptr<GemmTcPlan>> gemms_;
  std::vector<std::unique_ptr<VlaBlockPlan>> vla_plans_;
  // stand-alone cross-modal stage (hcm_run_cross_modal): its planned launches, and input-side plans re-encoded on the
  // caller's tensors (cached per distinct pointer set) so that no staging copy is needed
  struct CmStage {
    GemmTcPlan *insfc = nullptr, *visfc = nullptr, *kvx = nullptr;
    VlaBlockPlan* vla = nullptr;
    struct Entry {
      const void* key[3];
      std::unique_ptr<GemmTcPlan> insfc, visfc[2], kvx[2];
    };
    std::vector<Entry> cache;
  } cm_;

  // planned buffers
  h16 *tokens_r_ = nullptr, *cellmean_r_ = nullptr, *gmean_r_ = nullptr, *tokens_d_ = nullptr;
  h16* rgb_feat_ = nullptr;
  int rgb_fh_ = 0, rgb_fw_ = 0;
  float* gn_stats_arena_ = nullptr;
  size_t gn_stats_cap_ = 0, gn_stats_used_ = 0;
  h16 *bert_out_ = nullptr, *kvin_ = nullptr, *concat_hi_ = nullptr, *lo_in_ = nullptr, *vla_tokens_ = nullptr;
  float *gx_hi_ = nullptr, *y_hi_ = nullptr, *hscr_hi_ = nullptr, *gx_lo_ = nullptr, *y_lo_ = nullptr, *hscr_lo_ = nullptr;
  float *logits_buf_ = nullptr, *act_buf_ = nullptr, *stop_buf_ = nullptr, *hc_hi_buf_ = nullptr, *hc_lo_buf_ = nullptr;
  int64_t* subgoal_buf_ = nullptr;
  h16 *cm_bert_in_ = nullptr, *cm_kv_in_ = nullptr, *cm_out_ = nullptr;
  float *stage_rgb_ = nullptr, *stage_depth_ = nullptr, *stage_instr_ = nullptr, *stage_masks_ = nullptr,
        *stage_hc_hi_ = nullptr, *stage_hc_lo_ = nullptr;

  struct GraphEntry {
    const void* key[8];
    cudaGraphExec_t exec;
    int64_t launches;
    uint64_t last_use;
  };
  std::vector<GraphEntry> graphs_;
  struct HostGraphs {
    cudaGraphExec_t g1 = nullptr, g1b = nullptr, g2a = nullptr, g2b = nullptr;
    int64_t launches = 0;
    bool valid = false;
  } host_graphs_[2];   // per RGB input format
  int enc_mask_ = 7;             // run_encoders: 1 = RGB, 2 = depth, 4 = BERT branches (graph capture of subsets)
  uint64_t graph_tick_ = 0;
  int node_prio_[2] = {0, 0};    // per-launch priority attribute for depth / BERT ops (0 = none)
  int graph_miss_streak_ = 0;    // consecutive graph-cache misses (fresh pointers every call -> eager fallback)
  int eager_runs_ = 0;           // eager forward_policy calls since the last plan (kernels' one-time setup)
  void drop_graphs();
  // ROBOVLN_TIMELINE=path: eager multi-stream step with a timing event after every op of every stream; the
  // completion time of each op (relative to the step's start) is appended to `path` (diagnostics only)
  struct TlRec { int stream; const std::string* name; cudaEvent_t ev; };
  std::vector<TlRec> tl_;
  cudaEvent_t tl_start_ = nullptr;
  const char* tl_path_ = nullptr;
  void tl_mark(int stream_id, const std::string* name, cudaStream_t st);
  void tl_flush(cudaStream_t s);
  bool trunks_valid_ = false;
  bool streams_ready_ = false;
  cudaStream_t side_[2] = {nullptr, nullptr};
  cudaStream_t upload_ = nullptr;    // host entry: the RGB frame upload runs here, under the depth trunk and BERT
  cudaEvent_t ev_upload_ = nullptr;
  cudaStream_t aux_ = nullptr;       // host entry: depth-trunk graph
  cudaEvent_t ev_aux_[2] = {nullptr, nullptr};
  cudaStream_t capture_ = nullptr;   // graph capture happens here (the caller's stream may be the legacy default stream)
  cudaEvent_t events_[4] = {nullptr, nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> prof_events_;
};

}  // namespace rvb
