// C-ABI entry points (include/vqa_b200.h) and the whole-path orchestration.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace vqa {

// ---- thread-local error / launch accounting --------------------------------
static thread_local char g_error[512] = "";
static thread_local int g_launches = 0;

char* error_buffer() { return g_error; }
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}
void count_launch(int n) { g_launches += n; }
int launch_count() { return g_launches; }
void reset_launch_count() { g_launches = 0; }

struct DevInfo { int dev = -1, sms = 0, major = 0, minor = 0; };
static DevInfo query_device() {
  DevInfo d;
  if (cudaGetDevice(&d.dev) != cudaSuccess) { d.dev = -1; return d; }
  cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, d.dev);
  cudaDeviceGetAttribute(&d.major, cudaDevAttrComputeCapabilityMajor, d.dev);
  cudaDeviceGetAttribute(&d.minor, cudaDevAttrComputeCapabilityMinor, d.dev);
  return d;
}
static const DevInfo& device() {
  static thread_local DevInfo cache;
  int dev = -1;
  cudaGetDevice(&dev);
  if (cache.dev != dev || dev < 0) cache = query_device();
  return cache;
}
int sm_count() { const int n = device().sms; return n > 0 ? n : 148; }
int current_device() { return device().dev; }
int require_sm100() {
  const DevInfo& d = device();
  if (d.dev < 0) return fail(VQA_ERR_CUDA, "no CUDA device available (%s)", cudaGetErrorString(cudaGetLastError()));
  if (d.major != 10)
    return fail(VQA_ERR_UNSUPPORTED, "device compute capability %d.%d: this library is built for sm_100a only",
                d.major, d.minor);
  return VQA_OK;
}

// kernels implemented in the other translation units
int relation_labels(const float*, const float*, int, int, float, float, uint8_t*, cudaStream_t);
float relation_near_threshold(float, float);
int answer_scores(const int64_t*, const float*, int, int, int, float*, float*, float*, cudaStream_t);
int caption_gate_scale(const void*, const float*, const float*, int, int, int, int, void*, float*, cudaStream_t);
int seq_max(const void*, int, int, int, int, void*, cudaStream_t);
int softmax_mul(const float*, const void*, int, int, int, void*, cudaStream_t);
int attention_logits(const void*, int, const float*, int, const float*, int, int, int, int, int, float*, cudaStream_t);
int add_inplace(void*, const void*, size_t, int, cudaStream_t);
int attention_pool(const float*, int, float, const void*, int, int, int, int, float*, void*, void*, cudaStream_t);
int argmax_rows(const float*, int, int, int, int64_t*, cudaStream_t);
int split_f32(const float*, void*, void*, size_t, cudaStream_t);
int linear_tc_gru_step(const GruStepSplit&, cudaStream_t);
int gru_gate_table(const float*, const int64_t*, int, const float*, const float*, int, int, int, int, const float*, float*, void*,
                   cudaStream_t);
int embedding_gather(const int64_t*, int, int, int, int, const void*, void*, cudaStream_t, void* zero_ptr = nullptr,
                     size_t zero_bytes = 0);
int gru_gate(const float*, const float*, int, int, int, int, const float*, float*, void*, int, int, cudaStream_t);
int lstm_gate(const float*, int, int, float*, float*, void*, int, int, cudaStream_t);
int cast_f32_to_bf16(const float*, void*, size_t, cudaStream_t);
int cast_bf16_to_f32(const void*, float*, size_t, cudaStream_t);
int graph_attention(const vqa_graph_attention_args&, cudaStream_t);
int graph_attention_tc(const vqa_graph_attention_args&, cudaStream_t);
size_t optim_norm_workspace_bytes();
int grad_clip(const vqa_optim_tensor*, int, float, int, float*, float*, float*, cudaStream_t);
int adamax_step(const vqa_optim_tensor*, int, float, float, float, float, int, const float*, cudaStream_t);
size_t train_workspace_bytes(const vqa_train_args&);
int updown_train_step(const vqa_train_args&, cudaStream_t);
int gru_persistent(const void*, int, int, int, int, const void*, const void*, const float*, void*, int*, float*, void*,
                   void*, cudaStream_t);
int gru_pair(const void*, int, int, int, int, const void*, const void*, const float*, void*, int*, float*, void*,
             void*, const GruTrainSave*, int sm_limit, const GruTokenTable*, cudaStream_t);

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQA_B200_NO_PDL"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

bool l2_order_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQA_B200_NO_L2_ORDER"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

static bool force_simt() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQA_B200_FORCE_SIMT"); v = (e && e[0] == '1') ? 1 : 0; }
  return v == 1;
}

static int linear_dispatch(const vqa_linear_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.M >= 0 && a.N >= 1 && a.K >= 1, "vqa_linear: bad shape M=%d N=%d K=%d", a.M, a.N, a.K);
  if (a.M == 0) return VQA_OK;
  VQA_REQUIRE(a.d_A && a.d_W && a.d_out, "vqa_linear: NULL pointer");
  VQA_REQUIRE(a.dtype == VQA_F32 || a.dtype == VQA_BF16 || a.dtype == VQA_F16X2, "vqa_linear: dtype=%d", a.dtype);
  VQA_REQUIRE(a.out_dtype != VQA_F16X2 || a.dtype == VQA_F16X2, "vqa_linear: a split (f16x2) output needs split operands");
  VQA_REQUIRE(a.d_mul == nullptr || a.mul_row_div >= 1, "vqa_linear: mul_row_div must be >= 1");
  VQA_REQUIRE(a.d_add == nullptr || a.add_row_div >= 1, "vqa_linear: add_row_div must be >= 1");
  VQA_REQUIRE(!a.d_argmax_label || (a.d_argmax_ws && !a.d_logit_w && a.out_dtype == VQA_F32),
              "vqa_linear: fused argmax needs its workspace, the store form and an f32 output");
  if (a.dtype == VQA_F16X2) return linear_tc(a, s);                      // fp32-class: tensor cores or nothing
  if (a.dtype == VQA_BF16 && !force_simt()) return linear_tc(a, s);      // argmax in the epilogue
  if (int rc = linear_simt(a, s)) return rc;
  if (a.d_argmax_label) return argmax_rows((const float*)a.d_out, a.M, a.N, a.ldo, a.d_argmax_label, s);
  return VQA_OK;
}
static size_t argmax_ws_bytes(int M) { return align_up((size_t)(M > 0 ? M : 1) * 8, 256) + align_up((size_t)((M + 127) / 128 + 1) * 4, 256); }

static int part_width(int dtype) {
  return (dtype == VQA_F16X2 || (dtype == VQA_BF16 && !force_simt())) ? linear_tc_part_width() : 128;
}

// ---- GRU --------------------------------------------------------------------
struct GruWs { void* X; float* gi; float* gh; float* h; void* h_lp; void* h_op; int* counter; size_t bytes; };
static GruWs carve_gru(void* base, int B, int T, int H, int E_pad, int dtype) {
  GruWs w;
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t n) { void* r = p ? p + off : nullptr; off += align_up(n, 256); return r; };
  w.X = take((size_t)B * T * E_pad * elem_size(dtype));
  w.gi = (float*)take((size_t)B * T * 3 * H * 4);          // generic per-step path only
  w.gh = (float*)take((size_t)B * 3 * H * 4);
  w.h = (float*)take((size_t)B * H * 4);
  w.h_lp = take((size_t)B * H * elem_size(dtype));
  w.h_op = take((size_t)2 * B * H * 2);                    // persistent kernel: bf16 state x 2
  w.counter = (int*)take(256);
  w.bytes = off;
  return w;
}

// CTA pairs (tcgen05 cta_group::2, gru_pair.cu: 138 vs 144 us at B=1024, T=14) when the device can hold the pairs,
// else one CTA per tile; VQA_B200_GRU_PAIR=0 forces the single-CTA kernel.
// Nsight Compute cannot launch a cooperative CLUSTER kernel ("LaunchFailed" under the profiler and the target process is
// torn down): when its injection environment is present the single-CTA kernel runs instead, so `ncu python bench.py` works
// (VQA_B200_GRU_PAIR=1 overrides).
static bool gru_pair_enabled() {
  static int pair = -1;
  if (pair < 0) {
    const char* e = getenv("VQA_B200_GRU_PAIR");
    if (e) pair = (e[0] == '0') ? 0 : 1;
    else pair = getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ? 0 : 1;
  }
  return pair != 0;
}

// fp32-class mode (VQA_F16X2): token table for the input half (f32 [rows, 3H] = W_ih·emb[v] + b_ih), per step one split
// GEMM h·W_hhᵀ + b_hh (three tcgen05.mma per k-step) and the gate kernel, which writes the state as f32 and as the fp16
// plane pair that is the next GEMM's operand.  h_0 = 0: the first step has no GEMM.
static int gru_last_state_split(const vqa_gru_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.d_tokens && a.d_gi_table && a.d_w_hh && a.d_b_hh && a.d_h_last && !a.d_x && !a.d_out_all,
              "vqa_gru_last_state(f16x2): needs tokens, the f32 token table, W_hh planes, b_hh and d_h_last (no sequence form)");
  VQA_REQUIRE(a.B >= 0 && a.T >= 1 && a.H >= 1 && a.H % 8 == 0, "vqa_gru_last_state(f16x2): bad dims");
  const GruWs need = carve_gru(nullptr, a.B, a.T, a.H, a.E_pad, a.dtype);
  VQA_REQUIRE(a.d_workspace && a.workspace_bytes >= need.bytes,
              "vqa_gru_last_state: workspace %zu < %zu bytes", a.workspace_bytes, need.bytes);
  if (a.B == 0) return VQA_OK;
  const GruWs w = carve_gru(a.d_workspace, a.B, a.T, a.H, a.E_pad, a.dtype);
  void* tail = (char*)a.d_workspace + need.bytes;
  const size_t tail_bytes = (a.workspace_bytes - need.bytes) / 16 * 16;
  if (tail_bytes) VQA_CUDA_CHECK(cudaMemsetAsync(tail, 0, tail_bytes, s));
  int rc;
  // the state's plane pair ping-pongs between two buffers ([2][2][B][H]: h_lp and h_op are adjacent in the workspace): a
  // step's CTAs read ALL units of their rows as the A operand while other CTAs already write the new state of theirs
  char* planes4 = (char*)w.h_lp;
  const size_t pair_bytes = (size_t)a.B * a.H * 4;
  VQA_REQUIRE((char*)w.h_op == planes4 + pair_bytes, "vqa_gru_last_state(f16x2): workspace layout");
  auto planes_of = [&](int t) { return (void*)(planes4 + (size_t)(t & 1) * pair_bytes); };
  // step 0: h_0 = 0, no GEMM
  {
    const bool last = a.T == 1;
    if ((rc = gru_gate_table((const float*)a.d_gi_table, a.d_tokens, a.ntoken_rows, nullptr, a.d_b_hh, a.B, a.H, a.T, 0, nullptr,
                             last ? a.d_h_last : w.h, (last && a.d_h_last_lp) ? a.d_h_last_lp : planes_of(0), s))) return rc;
    if (last) return VQA_OK;
  }
  if (a.d_wh_packed != nullptr && a.H % 32 == 0) {
    // packed W_hh planes (32-unit blocks): GEMM + gates in one kernel, all steps in one launch when the tiles fit the device
    GruStepSplit g{};
    g.B = a.B; g.T = a.T; g.t = 1; g.t_end = a.T; g.H = a.H; g.ntoken_rows = a.ntoken_rows;
    g.tokens = a.d_tokens; g.gi_table = (const float*)a.d_gi_table; g.b_hh = a.d_b_hh;
    g.wh_packed_planes = a.d_wh_packed; g.h_planes = planes4; g.h_planes_last = a.d_h_last_lp;
    g.h = w.h; g.h_out_last = a.d_h_last; g.counter = w.counter;
    rc = linear_tc_gru_step(g, s);
    if (rc != VQA_ERR_UNSUPPORTED) return rc;
  }
  for (int t = 1; t < a.T; ++t) {                                   // no CTA pairs on this device: two kernels per step
    const bool last = (t == a.T - 1);
    vqa_linear_args gh{};
    gh.d_A = planes_of(t - 1); gh.lda = a.H; gh.d_W = a.d_w_hh; gh.ldw = a.H;
    gh.M = a.B; gh.N = 3 * a.H; gh.K = a.H; gh.dtype = VQA_F16X2;
    gh.d_bias = a.d_b_hh; gh.d_out = w.gh; gh.ldo = 3 * a.H; gh.out_dtype = VQA_F32; gh.mul_row_div = 1;
    if ((rc = linear_dispatch(gh, s))) return rc;
    if ((rc = gru_gate_table((const float*)a.d_gi_table, a.d_tokens, a.ntoken_rows, w.gh, a.d_b_hh, a.B, a.H, a.T, t, w.h,
                             last ? a.d_h_last : w.h, (last && a.d_h_last_lp) ? a.d_h_last_lp : planes_of(t), s))) return rc;
  }
  return VQA_OK;
}

// sm_limit > 0: the persistent kernel may use at most that many SMs (vqa_forward's overlap mode)
static int gru_last_state(const vqa_gru_args& a, cudaStream_t s, int sm_limit = 0) {
  if (a.dtype == VQA_F16X2) return gru_last_state_split(a, s);
  VQA_REQUIRE(((a.d_tokens && a.d_emb) || a.d_x) && a.d_w_ih && a.d_w_hh && a.d_b_ih && a.d_b_hh &&
              (a.d_h_last || a.d_out_all), "vqa_gru_last_state: NULL pointer");
  VQA_REQUIRE(a.B >= 0 && a.T >= 1 && a.H >= 1 && a.E_pad >= 1, "vqa_gru_last_state: bad dims");
  const GruWs need = carve_gru(nullptr, a.B, a.T, a.H, a.E_pad, a.dtype);
  VQA_REQUIRE(a.d_workspace && a.workspace_bytes >= need.bytes,
              "vqa_gru_last_state: workspace %zu < %zu bytes", a.workspace_bytes, need.bytes);
  if (a.B == 0) return VQA_OK;
  const GruWs w = carve_gru(a.d_workspace, a.B, a.T, a.H, a.E_pad, a.dtype);
  int rc;
  const void* X = a.d_x;
  // bytes of the workspace beyond what the GRU needs are zero-filled by the call (vqa_b200.h): by the gather kernel when
  // there is one, else by a memset
  void* tail = (char*)a.d_workspace + need.bytes;
  const size_t tail_bytes = (a.workspace_bytes - need.bytes) / 16 * 16;
  const bool fused_ok = a.dtype == VQA_BF16 && a.d_wx_packed && a.d_wh_packed && a.d_bias_packed && !force_simt() &&
                        a.H % 64 == 0 && a.E_pad % 64 == 0;
  // token-table form (d_gi_table): no gather, no x-part; pair kernel only (VQA_B200_GRU_TABLE=0 ignores the table)
  static int use_table = -1;
  if (use_table < 0) { const char* e = getenv("VQA_B200_GRU_TABLE"); use_table = (e && e[0] == '0') ? 0 : 1; }
  if (use_table && fused_ok && a.d_gi_table && a.d_tokens && !a.d_x && !a.d_out_all && gru_pair_enabled()) {
    // the counter block is the last 256 bytes of the GRU's own workspace: one memset clears it and the caller's tail
    const GruTokenTable tab{a.d_gi_table, a.d_tokens, a.ntoken_rows, tail_bytes};
    rc = gru_pair(nullptr, a.B, a.T, a.H, a.E_pad, a.d_wx_packed, a.d_wh_packed, a.d_bias_packed, w.h_op, w.counter,
                  a.d_h_last, a.d_h_last_lp, nullptr, nullptr, sm_limit, &tab, s);
    if (rc != VQA_ERR_UNSUPPORTED) return rc;
  }
  if (X && tail_bytes) VQA_CUDA_CHECK(cudaMemsetAsync(tail, 0, tail_bytes, s));
  if (!X) {
    if ((rc = embedding_gather(a.d_tokens, a.B * a.T, a.E_pad, a.ntoken_rows, a.dtype, a.d_emb, w.X, s, tail, tail_bytes)))
      return rc;
    X = w.X;
  }
  if (fused_ok) {
    if (gru_pair_enabled()) {
      rc = gru_pair(X, a.B, a.T, a.H, a.E_pad, a.d_wx_packed, a.d_wh_packed, a.d_bias_packed, w.h_op, w.counter,
                    a.d_h_last, a.d_h_last_lp, a.d_out_all, nullptr, sm_limit, nullptr, s);
      if (rc != VQA_ERR_UNSUPPORTED) return rc;
    }
    return gru_persistent(X, a.B, a.T, a.H, a.E_pad, a.d_wx_packed, a.d_wh_packed, a.d_bias_packed, w.h_op, w.counter,
                          a.d_h_last, a.d_h_last_lp, a.d_out_all, s);
  }
  // gi = X W_ihᵀ + b_ih for all T steps at once: [B*T, 3H] f32
  vqa_linear_args gi{};
  gi.d_A = X; gi.lda = a.E_pad; gi.d_W = a.d_w_ih; gi.ldw = a.E_pad;
  gi.M = a.B * a.T; gi.N = 3 * a.H; gi.K = a.E_pad; gi.dtype = a.dtype;
  gi.d_bias = a.d_b_ih; gi.d_out = w.gi; gi.ldo = 3 * a.H; gi.out_dtype = VQA_F32; gi.mul_row_div = 1;
  if ((rc = linear_dispatch(gi, s))) return rc;
  VQA_CUDA_CHECK(cudaMemsetAsync(w.h, 0, (size_t)a.B * a.H * 4, s));
  VQA_CUDA_CHECK(cudaMemsetAsync(w.h_lp, 0, (size_t)a.B * a.H * elem_size(a.dtype), s));
  const size_t es = elem_size(a.dtype);
  for (int t = 0; t < a.T; ++t) {
    // operand of this step's recurrent GEMM: the [B,H] copy, or (sequence form) the previous slice of [B,T,H]
    const bool from_all = a.d_out_all && t > 0;
    vqa_linear_args gh{};
    gh.d_A = from_all ? (const void*)((const char*)a.d_out_all + (size_t)(t - 1) * a.H * es) : w.h_lp;
    gh.lda = from_all ? a.T * a.H : a.H;
    gh.d_W = a.d_w_hh; gh.ldw = a.H;
    gh.M = a.B; gh.N = 3 * a.H; gh.K = a.H; gh.dtype = a.dtype;
    gh.d_bias = a.d_b_hh; gh.d_out = w.gh; gh.ldo = 3 * a.H; gh.out_dtype = VQA_F32; gh.mul_row_div = 1;
    if ((rc = linear_dispatch(gh, s))) return rc;
    const bool last = (t == a.T - 1);
    // state kept in f32 (w.h, updated in place) plus the low-precision copy that feeds the
    // next step's GEMM; the last step writes the caller's buffers
    void* lp = a.d_out_all ? (void*)((char*)a.d_out_all + (size_t)t * a.H * es)
                           : ((last && a.d_h_last_lp) ? a.d_h_last_lp : w.h_lp);
    if ((rc = gru_gate(w.gi, w.gh, a.B, a.H, a.T, t, w.h, (last && a.d_h_last) ? a.d_h_last : w.h, lp,
                       a.d_out_all ? a.T * a.H : a.H, a.dtype, s))) return rc;
  }
  return VQA_OK;
}

// ---- LSTM (rnn_type='LSTM' of SentenceEmbedding, modules.py:121-130): generic per-step path -----------------
// gi = X W_ihᵀ + b_ih for all T steps in one GEMM; per step ONE GEMM (h W_hhᵀ + b_hh + gi[:, t] as the additive
// epilogue operand) and the gate kernel.  h0 = c0 = 0 (modules.py:139-146).
struct LstmWs { float* gi; float* gates; float* c; void* h_lp; size_t bytes; };
static LstmWs carve_lstm(void* base, int B, int T, int H, int dtype) {
  LstmWs w; size_t off = 0; char* p = (char*)base;
  auto take = [&](size_t n) { void* r = p ? p + off : nullptr; off += align_up(n, 256); return r; };
  w.gi = (float*)take((size_t)B * T * 4 * H * 4);
  w.gates = (float*)take((size_t)B * 4 * H * 4);
  w.c = (float*)take((size_t)B * H * 4);
  w.h_lp = take((size_t)B * H * elem_size(dtype));
  w.bytes = off;
  return w;
}
static int lstm_sequence(const vqa_lstm_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.d_x && a.d_w_ih && a.d_w_hh && a.d_b_ih && a.d_b_hh && (a.d_h_last || a.d_out_all),
              "vqa_lstm_sequence: NULL pointer");
  VQA_REQUIRE(a.B >= 0 && a.T >= 1 && a.H >= 1 && a.E_pad >= 1, "vqa_lstm_sequence: bad dims");
  const LstmWs need = carve_lstm(nullptr, a.B, a.T, a.H, a.dtype);
  VQA_REQUIRE(a.d_workspace && a.workspace_bytes >= need.bytes, "vqa_lstm_sequence: workspace %zu < %zu bytes",
              a.workspace_bytes, need.bytes);
  if (a.B == 0) return VQA_OK;
  const LstmWs w = carve_lstm(a.d_workspace, a.B, a.T, a.H, a.dtype);
  const size_t es = elem_size(a.dtype);
  int rc;
  vqa_linear_args gi{};
  gi.d_A = a.d_x; gi.lda = a.E_pad; gi.d_W = a.d_w_ih; gi.ldw = a.E_pad;
  gi.M = a.B * a.T; gi.N = 4 * a.H; gi.K = a.E_pad; gi.dtype = a.dtype;
  gi.d_bias = a.d_b_ih; gi.d_out = w.gi; gi.ldo = 4 * a.H; gi.out_dtype = VQA_F32; gi.mul_row_div = 1; gi.add_row_div = 1;
  if ((rc = linear_dispatch(gi, s))) return rc;
  VQA_CUDA_CHECK(cudaMemsetAsync(w.c, 0, (size_t)a.B * a.H * 4, s));
  VQA_CUDA_CHECK(cudaMemsetAsync(w.h_lp, 0, (size_t)a.B * a.H * es, s));
  for (int t = 0; t < a.T; ++t) {
    const bool from_all = a.d_out_all && t > 0;
    vqa_linear_args gh{};
    gh.d_A = from_all ? (const void*)((const char*)a.d_out_all + (size_t)(t - 1) * a.H * es) : w.h_lp;
    gh.lda = from_all ? a.T * a.H : a.H;
    gh.d_W = a.d_w_hh; gh.ldw = a.H; gh.M = a.B; gh.N = 4 * a.H; gh.K = a.H; gh.dtype = a.dtype;
    gh.d_bias = a.d_b_hh; gh.d_add = w.gi + (size_t)t * 4 * a.H; gh.ld_add = a.T * 4 * a.H; gh.add_row_div = 1;
    gh.d_out = w.gates; gh.ldo = 4 * a.H; gh.out_dtype = VQA_F32; gh.mul_row_div = 1;
    if ((rc = linear_dispatch(gh, s))) return rc;
    const bool last = (t == a.T - 1);
    void* lp = a.d_out_all ? (void*)((char*)a.d_out_all + (size_t)t * a.H * es) : w.h_lp;
    if ((rc = lstm_gate(w.gates, a.B, a.H, w.c, last ? a.d_h_last : nullptr, lp, a.d_out_all ? a.T * a.H : a.H, a.dtype, s)))
      return rc;
  }
  return VQA_OK;
}

}  // namespace vqa

using namespace vqa;

// =============================================================================
extern "C" {

int vqa_abi_version(void) { return VQA_B200_ABI_VERSION; }
const char* vqa_last_error(void) { return error_buffer(); }

int vqa_device_info(int* sms, int* major, int* minor) {
  const DevInfo d = query_device();
  if (d.dev < 0) return fail(VQA_ERR_CUDA, "no CUDA device available");
  if (sms) *sms = d.sms;
  if (major) *major = d.major;
  if (minor) *minor = d.minor;
  return VQA_OK;
}

int vqa_relation_labels(const float* d_bbox, const float* d_wh, int B, int K, float img_w, float img_h,
                        uint8_t* d_labels, void* stream) {
  if (int rc = require_sm100()) return rc;
  return relation_labels(d_bbox, d_wh, B, K, img_w, img_h, d_labels, (cudaStream_t)stream);
}

float vqa_relation_near_threshold(float img_w, float img_h) { return relation_near_threshold(img_w, img_h); }

int vqa_relation_labels_host(const float* h_bbox, int B, int K, float img_w, float img_h, uint8_t* h_labels) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(h_bbox && h_labels && B >= 0 && K >= 1, "vqa_relation_labels_host: bad arguments");
  if (B == 0) return VQA_OK;
  float* d_bbox = nullptr;
  uint8_t* d_lab = nullptr;
  cudaStream_t s;
  VQA_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  int rc = VQA_OK;
  const size_t nb = (size_t)B * K * 4 * sizeof(float), nl = (size_t)B * K * K;
  cudaError_t e;
  if ((e = cudaMallocAsync((void**)&d_bbox, nb, s)) != cudaSuccess ||
      (e = cudaMallocAsync((void**)&d_lab, nl, s)) != cudaSuccess ||
      (e = cudaMemcpyAsync(d_bbox, h_bbox, nb, cudaMemcpyHostToDevice, s)) != cudaSuccess) {
    rc = fail(VQA_ERR_CUDA, "vqa_relation_labels_host: %s", cudaGetErrorString(e));
  }
  if (rc == VQA_OK) rc = relation_labels(d_bbox, nullptr, B, K, img_w, img_h, d_lab, s);
  if (rc == VQA_OK && (e = cudaMemcpyAsync(h_labels, d_lab, nl, cudaMemcpyDeviceToHost, s)) != cudaSuccess)
    rc = fail(VQA_ERR_CUDA, "vqa_relation_labels_host: %s", cudaGetErrorString(e));
  if (d_bbox) cudaFreeAsync(d_bbox, s);
  if (d_lab) cudaFreeAsync(d_lab, s);
  e = cudaStreamSynchronize(s);
  if (rc == VQA_OK && e != cudaSuccess) rc = fail(VQA_ERR_CUDA, "vqa_relation_labels_host: %s", cudaGetErrorString(e));
  cudaStreamDestroy(s);
  return rc;
}

int vqa_cast_f32_to_bf16(const float* d_src, void* d_dst, size_t n, void* stream) {
  if (int rc = require_sm100()) return rc;
  return cast_f32_to_bf16(d_src, d_dst, n, (cudaStream_t)stream);
}
int vqa_split_f32(const float* d_src, void* d_hi, void* d_lo, size_t n, void* stream) {
  if (int rc = require_sm100()) return rc;
  return split_f32(d_src, d_hi, d_lo, n, (cudaStream_t)stream);
}
int vqa_cast_bf16_to_f32(const void* d_src, float* d_dst, size_t n, void* stream) {
  if (int rc = require_sm100()) return rc;
  return cast_bf16_to_f32(d_src, d_dst, n, (cudaStream_t)stream);
}

int vqa_linear(const vqa_linear_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_linear: NULL args");
  return linear_dispatch(*args, (cudaStream_t)stream);
}
int vqa_linear_part_width(int dtype) { return part_width(dtype); }
int vqa_linear_tile_count(const vqa_linear_args* args) {
  if (!args || args->dtype != VQA_BF16 || force_simt()) return 0;
  return linear_tc_tile_count(*args);
}
int vqa_linear_tiles_n(const vqa_linear_args* args) {
  if (!args || args->dtype != VQA_BF16 || force_simt()) return 0;
  return linear_tc_tiles_n(*args);
}
size_t vqa_linear_argmax_workspace_bytes(int M) { return argmax_ws_bytes(M); }

size_t vqa_gru_workspace_bytes(int B, int T, int H, int E_pad, int dtype) {
  return carve_gru(nullptr, B, T, H, E_pad, dtype).bytes;
}
int vqa_gru_last_state(const vqa_gru_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_gru_last_state: NULL args");
  return gru_last_state(*args, (cudaStream_t)stream);
}

size_t vqa_lstm_workspace_bytes(int B, int T, int H, int dtype) { return carve_lstm(nullptr, B, T, H, dtype).bytes; }
int vqa_lstm_sequence(const vqa_lstm_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_lstm_sequence: NULL args");
  return lstm_sequence(*args, (cudaStream_t)stream);
}

int vqa_attention_pool(const float* d_logit_parts, int n_parts, float logit_bias, const void* d_x, int B,
                       int K, int V, int dtype, float* d_att, void* d_vsum, void* d_vatt, void* stream) {
  if (int rc = require_sm100()) return rc;
  return attention_pool(d_logit_parts, n_parts, logit_bias, d_x, B, K, V, dtype, d_att, d_vsum, d_vatt,
                        (cudaStream_t)stream);
}

int vqa_graph_attention(const vqa_graph_attention_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_graph_attention: NULL args");
  VQA_REQUIRE(args->layout == 0 || args->layout == 1, "vqa_graph_attention: layout=%d", args->layout);
  if (args->layout == 1) return graph_attention_tc(*args, (cudaStream_t)stream);
  return graph_attention(*args, (cudaStream_t)stream);
}

int vqa_argmax_rows(const float* d_logits, int B, int A, int ld, int64_t* d_label, void* stream) {
  if (int rc = require_sm100()) return rc;
  return argmax_rows(d_logits, B, A, ld, d_label, (cudaStream_t)stream);
}

int vqa_answer_scores(const int64_t* d_label, const float* d_target, int B, int A, int ld_target, float* d_scores_dense,
                      float* d_score_row, float* d_score_sum, void* stream) {
  if (int rc = require_sm100()) return rc;
  return answer_scores(d_label, d_target, B, A, ld_target, d_scores_dense, d_score_row, d_score_sum, (cudaStream_t)stream);
}

int vqa_caption_gate_scale(const void* d_out_w, const float* d_p, const float* d_r, int B, int T, int H, int dtype,
                           void* d_in2, float* d_a, void* stream) {
  if (int rc = require_sm100()) return rc;
  return caption_gate_scale(d_out_w, d_p, d_r, B, T, H, dtype, d_in2, d_a, (cudaStream_t)stream);
}
int vqa_seq_max(const void* d_e, int B, int T, int H, int dtype, void* d_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  return seq_max(d_e, B, T, H, dtype, d_out, (cudaStream_t)stream);
}
int vqa_softmax_mul(const float* d_z, const void* d_v, int B, int H, int dtype, void* d_out, void* stream) {
  if (int rc = require_sm100()) return rc;
  return softmax_mul(d_z, d_v, B, H, dtype, d_out, (cudaStream_t)stream);
}

int vqa_add_inplace(void* d_dst, const void* d_src, size_t n, int dtype, void* stream) {
  if (int rc = require_sm100()) return rc;
  return add_inplace(d_dst, d_src, n, dtype, (cudaStream_t)stream);
}
int vqa_attention_logits(const void* d_proj, int ldp, const float* d_q, int ldq, const float* d_w, int B, int K, int Hd,
                         int mode, int dtype, float* d_logits, void* stream) {
  if (int rc = require_sm100()) return rc;
  return attention_logits(d_proj, ldp, d_q, ldq, d_w, B, K, Hd, mode, dtype, d_logits, (cudaStream_t)stream);
}
int vqa_lstm_cell(const float* d_gates, int B, int H, int dtype, float* d_c, float* d_h_out, void* d_h_lp, int ld_lp,
                  void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(B >= 0 && H >= 1 && ld_lp >= H, "lstm_cell: bad dims B=%d H=%d ld_lp=%d", B, H, ld_lp);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(d_gates && d_c && d_h_lp, "lstm_cell: NULL pointer");
  return lstm_gate(d_gates, B, H, d_c, d_h_out, d_h_lp, ld_lp, dtype, (cudaStream_t)stream);
}
int vqa_gru_cell(const float* d_gi, const float* d_gh, const float* d_h_prev, int B, int H, int dtype, float* d_h_out,
                 void* d_h_lp, int ld_lp, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(B >= 0 && H >= 1 && ld_lp >= H, "gru_cell: bad dims B=%d H=%d ld_lp=%d", B, H, ld_lp);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(d_gi && d_gh && d_h_prev && d_h_out && d_h_lp, "gru_cell: NULL pointer");
  return gru_gate(d_gi, d_gh, B, H, 1, 0, d_h_prev, d_h_out, d_h_lp, ld_lp, dtype, (cudaStream_t)stream);
}

// Teacher-forced time loop of the caption head in one call (generator.py:99-111 around BaseDecoder.decode :168-181):
// per step W_q GEMM -> attention_logits -> attention_pool -> W_ih[:, E:] GEMM (+ the hoisted previous-word half) ->
// W_hh GEMM -> gate update; step t runs on the first batch_t samples (captions sorted by decreasing length).
struct DecWs { float* q; float* parts; void* att_v; float* gi; float* gh; size_t bytes; };
static DecWs carve_dec(char* base, int B, int K, int V, int Hd, int dtype, int ng = 4) {
  DecWs w{}; size_t off = 0;
  auto take = [&](size_t n) { char* p = base ? base + off : nullptr; off += align_up(n, 256); return p; };
  w.q = (float*)take((size_t)B * Hd * 4);
  w.parts = (float*)take((size_t)B * K * 4);
  w.att_v = take((size_t)B * V * elem_size(dtype));
  w.gi = (float*)take((size_t)B * ng * Hd * 4);           // sized for the LSTM cell's 4 gates (GRU uses 3)
  w.gh = (float*)take((size_t)B * ng * Hd * 4);
  w.bytes = off;
  return w;
}
size_t vqa_caption_decode_workspace_bytes(int B, int K, int V, int Hd, int dtype) {
  return carve_dec(nullptr, B, K, V, Hd, dtype).bytes;
}
int vqa_caption_decode_steps(const vqa_caption_decode_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_caption_decode_steps: NULL args");
  const vqa_caption_decode_args& a = *args;
  cudaStream_t s = (cudaStream_t)stream;
  VQA_REQUIRE(a.B >= 0 && a.T >= 0 && a.K >= 1 && a.V >= 8 && a.Hd >= 8 && a.Hd % 8 == 0,
              "caption_decode: bad dims B=%d T=%d K=%d V=%d Hd=%d", a.B, a.T, a.K, a.V, a.Hd);
  if (a.B == 0 || a.T == 0) return VQA_OK;
  VQA_REQUIRE(a.h_batches && a.d_x && a.d_proj && a.d_wq && a.d_logit_w && a.d_gi_prev && a.d_w_att && a.d_w_hh &&
              a.d_b_hh && a.d_h_all && a.d_h && a.d_h0_lp && a.d_workspace, "caption_decode: NULL pointer");
  VQA_REQUIRE(a.cell == 0 || (a.cell == 1 && a.d_c), "caption_decode: cell=%d (0 = GRUCell, 1 = LSTMCell with d_c)", a.cell);
  const int ng = a.cell == 1 ? 4 : 3;                      // gates per cell
  const DecWs w = carve_dec((char*)a.d_workspace, a.B, a.K, a.V, a.Hd, a.dtype);
  VQA_REQUIRE(a.workspace_bytes >= w.bytes, "caption_decode: workspace %zu < %zu bytes", a.workspace_bytes, w.bytes);
  const size_t es = elem_size(a.dtype);
  const void* h_in = a.d_h0_lp;
  size_t row = 0;
  int prev_bt = a.B, rc;
  for (int t = 0; t < a.T; ++t) {
    const int bt = a.h_batches[t];
    VQA_REQUIRE(bt >= 1 && bt <= prev_bt, "caption_decode: batch schedule must be non-increasing in [1,B] (step %d: %d)", t, bt);
    prev_bt = bt;
    vqa_linear_args q{};                       // hidden-state half of the attention
    q.d_A = h_in; q.lda = a.Hd; q.d_W = a.d_wq; q.ldw = a.Hd; q.M = bt; q.N = a.Hd; q.K = a.Hd; q.dtype = a.dtype;
    q.d_scale = a.d_wq_scale; q.d_bias = a.d_wq_bias; q.relu = a.att_mode == 0;
    q.d_out = w.q; q.ldo = a.Hd; q.out_dtype = VQA_F32; q.mul_row_div = 1; q.add_row_div = 1;
    if ((rc = linear_dispatch(q, s))) return rc;
    if ((rc = attention_logits(a.d_proj, a.Hd, w.q, a.Hd, a.d_logit_w, bt, a.K, a.Hd, a.att_mode, a.dtype, w.parts, s))) return rc;
    if ((rc = attention_pool(w.parts, 1, a.logit_bias, a.d_x, bt, a.K, a.V, a.dtype, nullptr, w.att_v, nullptr, s))) return rc;
    vqa_linear_args gi{};                      // W_ih[:, E:] att_v + (W_ih[:, :E] prev + b_ih)
    gi.d_A = w.att_v; gi.lda = a.V; gi.d_W = a.d_w_att; gi.ldw = a.V; gi.M = bt; gi.N = ng * a.Hd; gi.K = a.V; gi.dtype = a.dtype;
    gi.d_add = a.d_gi_prev + (size_t)t * ng * a.Hd; gi.ld_add = a.T * ng * a.Hd; gi.add_row_div = 1;
    gi.d_out = w.gi; gi.ldo = ng * a.Hd; gi.out_dtype = VQA_F32; gi.mul_row_div = 1;
    if ((rc = linear_dispatch(gi, s))) return rc;
    vqa_linear_args gh{};
    gh.d_A = h_in; gh.lda = a.Hd; gh.d_W = a.d_w_hh; gh.ldw = a.Hd; gh.M = bt; gh.N = ng * a.Hd; gh.K = a.Hd; gh.dtype = a.dtype;
    gh.d_bias = a.d_b_hh; gh.d_out = w.gh; gh.ldo = ng * a.Hd; gh.out_dtype = VQA_F32; gh.mul_row_div = 1; gh.add_row_div = 1;
    if (a.cell == 1) { gh.d_add = w.gi; gh.ld_add = ng * a.Hd; }        // LSTM: all four gate pre-activations summed here
    if ((rc = linear_dispatch(gh, s))) return rc;
    void* h_out_lp = (char*)a.d_h_all + row * a.Hd * es;
    if (a.cell == 1) {
      if ((rc = lstm_gate(w.gh, bt, a.Hd, a.d_c, a.d_h, h_out_lp, a.Hd, a.dtype, s))) return rc;
    } else if ((rc = gru_gate(w.gi, w.gh, bt, a.Hd, 1, 0, a.d_h, a.d_h, h_out_lp, a.Hd, a.dtype, s))) return rc;
    h_in = h_out_lp;
    row += (size_t)bt;
  }
  return VQA_OK;
}

size_t vqa_grad_clip_workspace_bytes(void) { return optim_norm_workspace_bytes(); }
int vqa_grad_clip(const vqa_optim_tensor* h_tensors, int n_tensors, float max_norm, int scale_in_place, void* d_workspace,
                  float* d_total_norm, float* d_scale, void* stream) {
  if (int rc = require_sm100()) return rc;
  return grad_clip(h_tensors, n_tensors, max_norm, scale_in_place, (float*)d_workspace, d_total_norm, d_scale,
                   (cudaStream_t)stream);
}
int vqa_adamax_step(const vqa_optim_tensor* h_tensors, int n_tensors, float beta1, float beta2, float eps, float weight_decay,
                    int step, const float* d_grad_scale, void* stream) {
  if (int rc = require_sm100()) return rc;
  return adamax_step(h_tensors, n_tensors, beta1, beta2, eps, weight_decay, step, d_grad_scale, (cudaStream_t)stream);
}

// ---- whole path --------------------------------------------------------------
// kernels of two streams can run at the same time only outside profilers that serialise launches
static bool concurrent_kernels_ok() {
  static int v = -1;
  if (v < 0) {
    const char* b = getenv("CUDA_LAUNCH_BLOCKING");
    v = (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || (b && b[0] == '1')) ? 0 : 1;
  }
  return v == 1;
}
// ReGAT: the graph attention runs beside the wide projection and chases it (vqa_forward_args.gat_chase_sms)
static bool chase_mode(const vqa_forward_args& a) {
  return a.relation && a.gat_chase_sms > 0 && a.d_Wg3 && a.dtype == VQA_BF16 && !force_simt() && a.K == 36 &&
         a.B >= 128 && concurrent_kernels_ok();
}
static bool overlap_mode(const vqa_forward_args& a) {
  return !chase_mode(a) && a.overlap && a.dtype == VQA_BF16 && !force_simt() && a.d_wx_packed && a.d_wh_packed && a.d_bias_packed &&
         a.B >= 512 && a.H % 64 == 0 && a.E_pad % 64 == 0 && a.H % 8 == 0;
}

struct FwdWs {
  void* gru; size_t gru_bytes;
  float* h; void* h_lp; float* qq; float* parts; void* vsum; void* Y; void* proj; void* joint; void* hid;
  float* vsum_f32;                           // f16x2 + relation: the fp32 graph attention's output before it is split
  int* progress;                             // row-block counters of the wide projection (chase mode)
  void* amax; size_t amax_bytes;             // fused answer selection of the last classifier layer
  size_t bytes;
};
static FwdWs carve_fwd(const vqa_forward_args& a, void* base) {
  FwdWs w{};
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t n) { void* r = p ? p + off : nullptr; off += align_up(n, 256); return r; };
  const size_t es = elem_size(a.dtype);
  // [GRU workspace | keys + counters of the fused answer selection]: the GRU call zero-fills whatever follows its own part
  w.gru_bytes = carve_gru(nullptr, a.B, a.T, a.H, a.E_pad, a.dtype).bytes;
  w.amax_bytes = argmax_ws_bytes(a.B);
  w.gru = take(w.gru_bytes + w.amax_bytes);
  w.amax = p ? (char*)w.gru + w.gru_bytes : nullptr;
  w.h = (float*)take((size_t)a.B * a.H * 4);
  w.h_lp = take((size_t)a.B * a.H * es);
  w.qq = (float*)take((size_t)a.B * 2 * a.H * 4);
  const int n_parts = (a.H + part_width(a.dtype) - 1) / part_width(a.dtype);
  w.parts = (float*)take((size_t)a.B * a.K * n_parts * 4);
  w.vsum = take((size_t)a.B * a.V * es);
  w.vsum_f32 = (a.relation && a.dtype == VQA_F16X2) ? (float*)take((size_t)a.B * a.V * 4) : nullptr;
  w.Y = a.relation ? take((size_t)a.B * a.K * (a.d_Wg3 ? 3 : 4) * a.V * es) : nullptr;
  w.proj = (!a.relation && overlap_mode(a)) ? take((size_t)a.B * a.K * a.H * es) : nullptr;   // stored W_v projection
  w.progress = chase_mode(a) ? (int*)take(((size_t)(a.B * a.K + 255) / 256 * 2 + 1) * 4) : nullptr;
  w.joint = take((size_t)a.B * a.H * es);
  w.hid = take((size_t)a.B * 2 * a.H * es);
  w.bytes = off;
  return w;
}

size_t vqa_forward_workspace_bytes(const vqa_forward_args* args) {
  if (!args) return 0;
  return carve_fwd(*args, nullptr).bytes;
}

static thread_local int g_last_forward_launches = 0;
int vqa_forward_last_launch_count(void) { return g_last_forward_launches; }

// side stream + fork / join events of the two-stream schedule, one set per (thread, device); created by the first
// uncaptured call (stream creation is not a capturable operation)
struct SideCtx { cudaStream_t s = nullptr; cudaEvent_t fork = nullptr, join = nullptr; bool tried = false; };
static SideCtx* side_ctx() {
  static thread_local SideCtx ctx[64];
  const int dev = current_device();
  if (dev < 0 || dev >= 64) return nullptr;
  SideCtx& c = ctx[dev];
  if (!c.tried) {
    c.tried = true;
    // highest priority: when SMs free up, the block scheduler places the (latency-bound) encoder's CTAs first
    int prio_lo = 0, prio_hi = 0;
    (void)cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c.s, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&c.join, cudaEventDisableTiming) != cudaSuccess) {
      (void)cudaGetLastError();
      c.s = nullptr;
    }
  }
  return c.s ? &c : nullptr;
}

// Share (per mille) of the projection's tiles that the side SMs take once the question encoder is done: both sets of
// SMs should finish together.  tile time ~ 12.5 us per 256 x 256 x 2048 pair tile; encoder beside a running GEMM ~ 15.7 us
// per GRU step (two interleaved row blocks on 64 SMs) + 70 us (fork, gather, [W_q;q_net] on 64 SMs) — measured at H = 1024:
// 300 per mille is the best split of the wide ReGAT projection at B = 1024 (profiles/r02_overlap_split.md); the engine
// can override the estimate (side_tile_permille).
static int auto_side_permille(int tiles, int K, int T, int main_sms, int side_sms, bool token_table) {
  // (the token-table GRU saves the gather launch; beside a running GEMM its steps take as long as the x-part form's —
  // 216 us at T = 14 in both timelines — because the step is bound by the L2 -> SM stream it shares with the GEMM)
  const double tau = 12.5 * K / 2048.0, G = 15.7 * T + (token_table ? 64.0 : 70.0);
  const double pt = main_sms / 2, ps = side_sms / 2;
  const double t_end = (tiles * tau + ps * G) / (pt + ps);
  double share = ps * (t_end - G) / tau / tiles;
  if (share < 0.02) return 0;
  if (share > 0.5) share = 0.5;
  return (int)(share * 1000.0);
}

int vqa_forward(const vqa_forward_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_forward: NULL args");
  const vqa_forward_args& a = *args;
  cudaStream_t s = (cudaStream_t)stream;
  VQA_REQUIRE(a.B >= 0 && a.K >= 1 && a.V >= 1 && a.H >= 1 && a.A >= 1 && a.T >= 1, "vqa_forward: bad dims");
  VQA_REQUIRE(a.d_img && a.d_tokens && a.d_logits, "vqa_forward: NULL input/output");
  VQA_REQUIRE(a.d_emb && a.d_w_ih && a.d_w_hh && a.d_Wv && a.d_Wqq && a.d_wlin && a.d_Wvn && a.d_Wc0 && a.d_Wc1,
              "vqa_forward: NULL weight");
  const FwdWs need = carve_fwd(a, nullptr);
  VQA_REQUIRE(a.d_workspace && a.workspace_bytes >= need.bytes, "vqa_forward: workspace %zu < %zu bytes",
              a.workspace_bytes, need.bytes);
  if (a.B == 0) return VQA_OK;
  const FwdWs w = carve_fwd(a, a.d_workspace);
  const int before = launch_count();
  int rc;
  const uint8_t* labels = a.d_labels;
  if (a.relation) {
    if (a.d_Wg3) {
      VQA_REQUIRE(a.dtype == VQA_BF16 && a.d_wvec && a.d_label_bias_lp, "vqa_forward: merged ReGAT weights need bf16");
    } else {
      VQA_REQUIRE(a.d_Wg && a.d_label_bias && a.d_ba && a.d_bb, "vqa_forward: NULL ReGAT weight");
    }
    if (a.d_bbox != nullptr) {
      VQA_REQUIRE(a.d_labels_out, "vqa_forward: d_bbox given without d_labels_out");
      if ((rc = relation_labels(a.d_bbox, nullptr, a.B, a.K, a.img_w, a.img_h, a.d_labels_out, s))) return rc;
      labels = a.d_labels_out;
    }
    VQA_REQUIRE(labels, "vqa_forward: relation path needs d_labels or d_bbox");
    VQA_REQUIRE(a.d_att, "vqa_forward: relation path needs d_att");
  }
  if (a.att_concat) VQA_REQUIRE(a.d_W1q && a.d_b1, "vqa_forward: att_concat needs d_W1q and d_b1");
  if (a.dtype == VQA_F16X2)
    VQA_REQUIRE(a.d_gi_table && !a.d_Wg3 && a.V % 8 == 0 && a.H % 8 == 0 && (a.B * a.K * a.V) % 8 == 0,
                "vqa_forward(f16x2): needs the f32 token table, 16-byte aligned planes and the unmerged ReGAT weights");

  // ---- schedule: everything on `s` in order, or (overlap) question encoder on the side stream `sq` while the
  //      question-independent projection of the region features runs on `s` on the other SMs
  const bool chase = chase_mode(a);
  SideCtx* side = (overlap_mode(a) || chase) ? side_ctx() : nullptr;
  const bool overlap = side != nullptr && !chase;
  VQA_REQUIRE(side || !(overlap_mode(a) || chase), "vqa_forward: a two-stream schedule was requested but the side stream could not be created");
  const int sms = sm_count();
  int side_sms = overlap ? ((a.side_sms > 0 ? a.side_sms : 64) & ~1) : 0;
  if (overlap && (side_sms < 2 || side_sms > sms - 2)) side_sms = (sms / 2) & ~1;
  const int main_sms = (sms - side_sms) & ~1;
  cudaStream_t sq = overlap ? side->s : s;
  if (overlap) {
    VQA_CUDA_CHECK(cudaEventRecord(side->fork, s));
    VQA_CUDA_CHECK(cudaStreamWaitEvent(sq, side->fork, 0));
  }
  const int pw = overlap && !a.relation ? a.H : part_width(a.dtype);   // unfused logits come as ONE part per region
  const int n_parts = (a.H + pw - 1) / pw;
  const int maps = a.d_Wg3 ? 3 : 4;

  // the question-independent projection: the wide ReGAT GEMM on the raw features (x = a·v enters the graph attention
  // through its coefficients, graph_attn_tc.cu), or — overlap only — the stored W_v projection of the Up-Down path
  vqa_linear_args proj{};
  if (a.relation) {
    proj.d_A = a.d_img; proj.lda = a.V; proj.d_W = a.d_Wg3 ? a.d_Wg3 : a.d_Wg; proj.ldw = a.V; proj.M = a.B * a.K;
    proj.N = maps * a.V; proj.K = a.V; proj.dtype = a.dtype; proj.mul_row_div = 1; proj.d_out = w.Y; proj.ldo = maps * a.V;
    proj.out_dtype = a.dtype == VQA_F16X2 ? VQA_F32 : a.dtype;        // fp32-class: Y feeds the fp32 graph attention
  } else if (overlap) {
    // 'new' attention: ReLU(s·W_v v + b) (attention.py:70); 'base': the bare v-half s·W1v v of the concat layer — its bias,
    // the q-half and the ReLU follow in the logit kernel (attention.py:38-40)
    proj.d_A = a.d_img; proj.lda = a.V; proj.d_W = a.d_Wv; proj.ldw = a.V; proj.M = a.B * a.K; proj.N = a.H; proj.K = a.V;
    proj.dtype = a.dtype; proj.d_scale = a.d_sv; proj.mul_row_div = 1;
    if (!a.att_concat) { proj.d_bias = a.d_bv; proj.relu = 1; }
    proj.d_out = w.proj; proj.ldo = a.H; proj.out_dtype = a.dtype;
  }
  int split_tile = 0;                            // tiles [split_tile, end) go to the side SMs after the encoder
  if (overlap) {
    const int tiles = linear_tc_tile_count(proj);
    int permille = a.side_tile_permille;
    if (permille == 0) permille = auto_side_permille(tiles, a.V, a.T, main_sms, side_sms, a.d_gi_table != nullptr);
    if (permille < 0) permille = 0;
    if (permille > 900) permille = 900;
    split_tile = tiles - (int)((long long)tiles * permille / 1000);
    vqa_linear_args main_part = proj;
    main_part.tile_begin = 0; main_part.tile_end = split_tile; main_part.cta_limit = main_sms;
    if ((rc = linear_dispatch(main_part, s))) return rc;
  }

  // 1. question encoder (encoder.py:159-160)
  vqa_gru_args g{};
  g.d_tokens = a.d_tokens; g.B = a.B; g.T = a.T; g.H = a.H; g.E_pad = a.E_pad; g.ntoken_rows = a.ntoken_rows;
  g.dtype = a.dtype; g.d_emb = a.d_emb; g.d_w_ih = a.d_w_ih; g.d_b_ih = a.d_b_ih; g.d_w_hh = a.d_w_hh;
  g.d_b_hh = a.d_b_hh; g.d_wx_packed = a.d_wx_packed; g.d_wh_packed = a.d_wh_packed; g.d_bias_packed = a.d_bias_packed;
  g.d_workspace = w.gru; g.workspace_bytes = w.gru_bytes + w.amax_bytes;     // the tail (w.amax) comes back zeroed
  g.d_h_last = w.h; g.d_h_last_lp = w.h_lp; g.d_gi_table = a.d_gi_table;
  if ((rc = gru_last_state(g, sq, side_sms))) return rc;
  // 2. [W_q ; q_net] (attention.py:71, encoder.py:169): qq = ReLU(h Wqqᵀ s + b) f32 [B,2H]
  vqa_linear_args l{};
  if (!a.att_concat) {
    l.d_A = w.h_lp; l.lda = a.H; l.d_W = a.d_Wqq; l.ldw = a.H; l.M = a.B; l.N = 2 * a.H; l.K = a.H; l.dtype = a.dtype;
    l.d_scale = a.d_sqq; l.d_bias = a.d_bqq; l.relu = 1; l.mul_row_div = 1;
    l.d_out = w.qq; l.ldo = 2 * a.H; l.out_dtype = VQA_F32; l.cta_limit = side_sms;
    if ((rc = linear_dispatch(l, sq))) return rc;
  } else {
    // ConcatAttention (attention.py:38-42): W1[v;q] = W1v v + W1q q, so the q-half is one [B,H] GEMM
    // (no ReLU) that enters the W_v GEMM as an additive row-broadcast operand; q_net separately.
    l.d_A = w.h_lp; l.lda = a.H; l.d_W = a.d_W1q; l.ldw = a.H; l.M = a.B; l.N = a.H; l.K = a.H; l.dtype = a.dtype;
    l.d_scale = a.d_sv; l.d_bias = a.d_b1; l.relu = 0; l.mul_row_div = 1;
    l.d_out = w.qq; l.ldo = 2 * a.H; l.out_dtype = VQA_F32; l.cta_limit = side_sms;
    if ((rc = linear_dispatch(l, sq))) return rc;
    l = vqa_linear_args{};
    l.d_A = w.h_lp; l.lda = a.H; l.d_W = a.d_Wqq; l.ldw = a.H; l.M = a.B; l.N = a.H; l.K = a.H; l.dtype = a.dtype;
    l.d_scale = a.d_sqq; l.d_bias = a.d_bqq; l.relu = 1; l.mul_row_div = 1;
    l.d_out = w.qq + a.H; l.ldo = 2 * a.H; l.out_dtype = VQA_F32; l.cta_limit = side_sms;
    if ((rc = linear_dispatch(l, sq))) return rc;
  }
  if (overlap) {
    // the encoder's SMs take the tail of the projection, then the streams join
    vqa_linear_args side_part = proj;
    side_part.tile_begin = split_tile; side_part.tile_end = 0; side_part.cta_limit = side_sms;
    if ((rc = linear_dispatch(side_part, sq))) return rc;
    VQA_CUDA_CHECK(cudaEventRecord(side->join, sq));
    VQA_CUDA_CHECK(cudaStreamWaitEvent(s, side->join, 0));
  }
  // 3. attention logits (attention.py:70-75 / :38-42)
  if (overlap && !a.relation) {
    // the projection is stored: reduce it against the question half, one warp per region row
    if ((rc = attention_logits(w.proj, a.H, w.qq, 2 * a.H, a.d_wlin, a.B, a.K, a.H, a.att_concat ? 1 : 0, a.dtype, w.parts, s)))
      return rc;
  } else {
    // W_v projection fused with ⊙Qp (or + the q-half) and the 1-wide logit layer: the [B·K,H] projection never
    // reaches HBM
    l = vqa_linear_args{};
    l.d_A = a.d_img; l.lda = a.V; l.d_W = a.d_Wv; l.ldw = a.V; l.M = a.B * a.K; l.N = a.H; l.K = a.V; l.dtype = a.dtype;
    l.d_scale = a.d_sv; l.relu = 1; l.mul_row_div = 1;
    if (!a.att_concat) { l.d_bias = a.d_bv; l.d_mul = w.qq; l.ld_mul = 2 * a.H; l.mul_row_div = a.K; }
    else { l.d_add = w.qq; l.ld_add = 2 * a.H; l.add_row_div = a.K; }
    l.d_logit_w = a.d_wlin; l.d_out = w.parts; l.ldo = n_parts; l.out_dtype = VQA_F32;
    if ((rc = linear_dispatch(l, s))) return rc;
  }
  // 4. softmax over K + weighted sum (attention.py:86, encoder.py:166, predictor.py:85)
  if (!a.relation) {
    if ((rc = attention_pool(w.parts, n_parts, a.b_lin, a.d_img, a.B, a.K, a.V, a.dtype, a.d_att ? a.d_att : nullptr,
                             w.vsum, a.d_v, s))) return rc;
  } else {
    float* att = a.d_att;
    if ((rc = attention_pool(w.parts, n_parts, a.b_lin, a.d_img, a.B, a.K, a.V, a.dtype, att, nullptr, nullptr, s))) return rc;
    // 5. wide projection of the raw features (already done in overlap mode) + relation-masked graph attention (gcn.py)
    cudaStream_t sg = s;                         // stream of the graph attention
    int chase_sms = 0;
    if (chase) {
      // the graph attention chases the projection: GEMM on `s` (all SMs but chase_sms) publishes finished row blocks,
      // the graph attention on the side stream consumes them while they are still in L2
      chase_sms = a.gat_chase_sms < sms - 8 ? a.gat_chase_sms : sms - 8;
      VQA_CUDA_CHECK(cudaMemsetAsync(w.progress, 0, ((size_t)(a.B * a.K + 255) / 256 * 2 + 1) * 4, s));
      VQA_CUDA_CHECK(cudaEventRecord(side->fork, s));
      VQA_CUDA_CHECK(cudaStreamWaitEvent(side->s, side->fork, 0));
      proj.d_progress = w.progress;
      proj.cta_limit = (sms - chase_sms) & ~1;
      sg = side->s;
    }
    if (!overlap && (rc = linear_dispatch(proj, s))) return rc;
    vqa_graph_attention_args ga{};
    if (chase) { ga.d_progress = w.progress; ga.progress_target = linear_tc_tiles_n(proj); ga.cta_limit = chase_sms; }
    ga.d_Y = w.Y; ga.ldy = maps * a.V; ga.d_att = att; ga.d_labels = labels; ga.d_label_bias = a.d_label_bias;
    ga.num_labels = a.num_labels; ga.d_ba = a.d_ba; ga.d_bb = a.d_bb; ga.B = a.B; ga.K = a.K; ga.V = a.V; ga.dtype = a.dtype;
    ga.d_out = a.d_v; ga.d_vsum = w.vsum; ga.d_alpha = a.d_alpha;
    if (a.dtype == VQA_F16X2) {
      VQA_REQUIRE(!a.d_v, "vqa_forward(f16x2): the encoder output 'v' is not produced in this mode");
      ga.dtype = VQA_F32; ga.d_vsum = w.vsum_f32;
    }
    if (a.d_Wg3) {
      ga.layout = 1; ga.d_x = a.d_img; ga.ldx = a.V; ga.d_wvec = a.d_wvec; ga.c0 = a.gat_c0;
      ga.d_label_bias_lp = a.d_label_bias_lp;
      if ((rc = graph_attention_tc(ga, sg))) return rc;
    } else if ((rc = graph_attention(ga, s))) return rc;
    if (chase) {
      VQA_CUDA_CHECK(cudaEventRecord(side->join, sg));
      VQA_CUDA_CHECK(cudaStreamWaitEvent(s, side->join, 0));
    }
    if (a.dtype == VQA_F16X2 &&
        (rc = split_f32(w.vsum_f32, w.vsum, (char*)w.vsum + (size_t)a.B * a.V * 2, (size_t)a.B * a.V, s))) return rc;
  }
  // 6. v_net ⊙ q_net (predictor.py:88-91)
  l = vqa_linear_args{};
  l.d_A = w.vsum; l.lda = a.V; l.d_W = a.d_Wvn; l.ldw = a.V; l.M = a.B; l.N = a.H; l.K = a.V; l.dtype = a.dtype;
  l.d_scale = a.d_svn; l.d_bias = a.d_bvn; l.relu = 1; l.d_mul = w.qq + a.H; l.ld_mul = 2 * a.H; l.mul_row_div = 1;
  l.d_out = w.joint; l.ldo = a.H; l.out_dtype = a.dtype;
  if ((rc = linear_dispatch(l, s))) return rc;
  // 7. classifier (predictor.py:93; FCNet 2 layers, final ReLU modules.py:55)
  l = vqa_linear_args{};
  l.d_A = w.joint; l.lda = a.H; l.d_W = a.d_Wc0; l.ldw = a.H; l.M = a.B; l.N = 2 * a.H; l.K = a.H; l.dtype = a.dtype;
  l.d_scale = a.d_sc0; l.d_bias = a.d_bc0; l.relu = 1; l.mul_row_div = 1;
  l.d_out = w.hid; l.ldo = 2 * a.H; l.out_dtype = a.dtype;
  if ((rc = linear_dispatch(l, s))) return rc;
  l = vqa_linear_args{};
  l.d_A = w.hid; l.lda = 2 * a.H; l.d_W = a.d_Wc1; l.ldw = 2 * a.H; l.M = a.B; l.N = a.A; l.K = 2 * a.H; l.dtype = a.dtype;
  l.d_scale = a.d_sc1; l.d_bias = a.d_bc1; l.relu = 1; l.mul_row_div = 1;
  l.d_out = a.d_logits; l.ldo = a.A; l.out_dtype = VQA_F32;
  // 8. answers (wrapper.py:14): selected in this GEMM's epilogue (bf16 path) or by argmax_rows after it (fp32 path)
  l.d_argmax_label = a.d_label; l.d_argmax_ws = a.d_label ? w.amax : nullptr;
  if ((rc = linear_dispatch(l, s))) return rc;
  if (a.d_q) VQA_CUDA_CHECK(cudaMemcpy2DAsync(a.d_q, (size_t)a.H * 4, w.qq + a.H, (size_t)2 * a.H * 4, (size_t)a.H * 4, a.B,
                                              cudaMemcpyDeviceToDevice, s));
  g_last_forward_launches = launch_count() - before;
  return VQA_OK;
}

size_t vqa_train_workspace_bytes(const vqa_train_args* args) { return args ? train_workspace_bytes(*args) : 0; }
int vqa_updown_train_step(const vqa_train_args* args, void* stream) {
  if (int rc = require_sm100()) return rc;
  VQA_REQUIRE(args, "vqa_updown_train_step: NULL args");
  return updown_train_step(*args, (cudaStream_t)stream);
}

}  // extern "C"
