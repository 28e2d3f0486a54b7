/*
 * vqa_b200.h — C ABI of the B200-native VQA forward hot path.
 *
 * Drop-in boundary for Jayie/vqa-collection's Up-Down / ReGAT forward path.
 * The reference is pure Python/PyTorch (no FFI of its own); every entry point
 * below names the reference symbol (file:line, relative to the reference
 * repository root) whose arithmetic it replaces.  The Python host side in
 * vqa_collection_b200/ binds these with ctypes (see INTEGRATION.md for the stub
 * a reference maintainer would add).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary
 *   - pointers named d_* are DEVICE pointers (current device), h_* are HOST
 *     pointers (pinned memory recommended); `stream` is a cudaStream_t passed
 *     as void* (NULL = legacy default stream)
 *   - every function returns 0 on success, a negative vqa_status on failure and
 *     never falls back to a CPU path; vqa_last_error() returns the message of
 *     the last failure on the calling thread
 *   - all launches are asynchronous on `stream` unless the name ends in _host
 *   - matrices are row-major; "ld" is the leading dimension in ELEMENTS
 *   - sm_100a only: functions return VQA_ERR_UNSUPPORTED on other devices
 */
#ifndef VQA_B200_H_
#define VQA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VQA_B200_ABI_VERSION 11

typedef enum {
  VQA_OK = 0,
  VQA_ERR_INVALID = -1,      /* bad argument (shape, alignment, NULL)       */
  VQA_ERR_CUDA = -2,         /* CUDA runtime / driver error                 */
  VQA_ERR_UNSUPPORTED = -3   /* not an sm_100 device / unsupported config   */
} vqa_status;

typedef enum {
  VQA_F32 = 0,               /* fp32 operands, fp32 FFMA accumulate          */
  VQA_BF16 = 1,              /* bf16 operands, fp32 accumulate (tcgen05/TMEM)*/
  /* fp32-class on tensor cores: a tensor of R rows and leading dimension ld is a PAIR OF fp16 PLANES [2][R][ld],
   * x = hi + lo'·2^-11 with hi = fp16(x), lo' = fp16((x - hi)·2^11) (22 significant bits; vqa_split_f32 makes them).
   * vqa_linear then issues three tcgen05.mma per k-step (hi·hi + 2^-11·(hi·lo' + lo'·hi), fp32 accumulators in TMEM). */
  VQA_F16X2 = 2
} vqa_dtype;

int vqa_abi_version(void);
const char* vqa_last_error(void);
/* sm count / compute capability of the current device */
int vqa_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------
 * k7  spatial relation labels
 * replaces util/relation.py:65-80 relation_graph (and :3-45 spatial_relation),
 * batched: bbox [B,K,4] f32 (x0,y0,x1,y1) -> labels [B,K,K] u8 in {0..11}.
 * One evaluation per unordered pair i<j writes [i,j] and [j,i]; diagonal 0.
 * Image size: per image from d_wh [B,2] f32 (w,h) when non-NULL, else the
 * uniform (img_w, img_h).  K <= 64.
 * ---------------------------------------------------------------------- */
int vqa_relation_labels(const float* d_bbox, const float* d_wh, int B, int K, float img_w,
                        float img_h, uint8_t* d_labels, void* stream);
/* Host-only helper (no GPU needed): the float32 threshold s_max the kernel compares the SQUARED
 * centre distance against — the largest float32 s whose correctly rounded sqrt, widened to
 * double, is <= ||(img_w,img_h)||/2, i.e. exactly relation.py:37-38's `dist <= 0.5`.  Exposed so
 * the sqrt/divide-free form can be pinned against numpy on the CPU. */
float vqa_relation_near_threshold(float img_w, float img_h);
/* host-buffer form (H2D + kernel + D2H + sync), the e2e path of the label op */
int vqa_relation_labels_host(const float* h_bbox, int B, int K, float img_w, float img_h,
                             uint8_t* h_labels);

/* ------------------------------------------------------------------------
 * element type conversion (wire format f32 -> resident bf16), n elements
 * ---------------------------------------------------------------------- */
int vqa_cast_f32_to_bf16(const float* d_src, void* d_dst, size_t n, void* stream);
int vqa_cast_bf16_to_f32(const void* d_src, float* d_dst, size_t n, void* stream);
/* f32 -> the fp16 plane pair of VQA_F16X2 (hi, lo' = residual·2^11): n contiguous elements into each plane */
int vqa_split_f32(const float* d_src, void* d_hi, void* d_lo, size_t n, void* stream);

/* ------------------------------------------------------------------------
 * k4/k5/k8/k11  fused weight-normed linear layer
 * replaces modules/modules.py:13-60 FCNet (one weight_norm(nn.Linear)+ReLU
 * stage), attention.py:66,75 (the 1-wide logit layer, fused as a row
 * reduction) and the plain nn.Linear maps of gcn.py:101-103 / modules.py:92-93.
 *
 *   y[m,n] = (sum_k A[m,k] * W[n,k]) * scale[n] + bias[n]
 *   if add && !add_after_act:  y += add[(m / add_row_div), n]   (f32, ld_add; the q-half of
 *             ConcatAttention's first layer, attention.py:38-42, broadcast over K)
 *   if relu:  y = y > 0 ? y : leaky_slope * y         (leaky_slope = 0: ReLU; != 0: the LeakyReLU of
 *             modules.py:62-78 LReLUNet)
 *   if add && add_after_act:  y += add[...]           (predictor.py:209  LReLU(W v) + c)
 *   if mul:   y *= mul[(m / mul_row_div), n]          (f32, ld_mul)
 *   if sigmoid: y = 1 / (1 + exp(-y))                 (predictor.py:181-184 classifier)
 *   if logit_w == NULL:  out[m,n] = y                 (out_dtype, ldo)
 *   else: out_f32[m * n_parts + p] = sum_{n in part p} y * logit_w[n]
 *         with n_parts = ceil(N / vqa_linear_part_width(dtype))
 *
 * dtype VQA_BF16: A, W are bf16, lda/ldw % 8 == 0, 16-byte aligned bases (TMA; a K
 * tail is zero-filled); tcgen05.mma 128xBN tiles with TMEM accumulators.
 * dtype VQA_F32 : A, W are f32, K % 16 == 0; FFMA tiles.
 * ---------------------------------------------------------------------- */
typedef struct {
  const void* d_A;  int lda;
  const void* d_W;  int ldw;
  int M, N, K;
  int dtype;                 /* vqa_dtype of A and W                         */
  const float* d_scale;      /* [N] or NULL (=1)                             */
  const float* d_bias;       /* [N] or NULL (=0)                             */
  int relu;
  const float* d_mul;        /* optional elementwise multiplier, or NULL     */
  int ld_mul;
  int mul_row_div;           /* row m reads mul row m / mul_row_div (>=1)    */
  const float* d_logit_w;    /* [N] or NULL; selects the row-reduction form  */
  void* d_out;               /* [M,ldo] out_dtype, or [M,n_parts] f32        */
  int ldo;
  int out_dtype;             /* vqa_dtype of out (ignored in logit form)     */
  const float* d_add;        /* optional additive row-broadcast operand      */
  int ld_add;
  int add_row_div;           /* row m reads add row m / add_row_div (>=1)    */
  /* training-step forms (backward GEMMs on the same row-major tensors, no transposes):
   *   trans_w: W is given as [K, N] row-major (ldw): y = A * W        e.g. dX = dY * W_fwd
   *   trans_a: A is given as [K, M] row-major (lda): y = A^T * W      e.g. dW = dY^T * X
   *            (trans_a needs trans_w).  bf16: K need not be a multiple of 64 (TMA zero fill).
   *   mask:    y = (mask[m,n] > 0) ? y : 0, applied last (after relu and mul): the backward
   *            of a ReLU whose saved OUTPUT is `mask` (mask_dtype, ld_mask)               */
  int trans_a, trans_w;
  const void* d_mask; int ld_mask; int mask_dtype;
  /* config-5 epilogue forms (all 0 = the behaviour above) */
  float leaky_slope;         /* negative-side slope of the activation when relu != 0 */
  int add_after_act;         /* apply `add` after the activation instead of before   */
  int sigmoid;               /* logistic function applied last                       */
  /* fused answer selection (wrapper.py:14, torch.max(predict, 1)[1]: lowest index among equal maxima) over the N
   * outputs of each row, store form only: d_argmax_label int64 [M]; d_argmax_ws = vqa_linear_argmax_workspace_bytes(M)
   * bytes that are ZERO on entry (the call leaves them dirty).  NULL = off. */
  int64_t* d_argmax_label;
  void* d_argmax_ws;
  /* launch shaping (bf16 tensor-core path; all 0 = the whole GEMM on the whole device): the call computes only the
   * output tiles [tile_begin, tile_end) of the kernel's tile order (vqa_linear_tile_count gives the total) on at most
   * cta_limit SMs, so that two calls on different streams can split one GEMM between two sets of SMs. */
  int tile_begin, tile_end;
  int cta_limit;
  /* optional progress counters int32 [2*ceil(M/256)], ZERO on entry (bf16 tensor-core path, store form): every finished
   * output tile adds 1 (release, gpu scope) to the counter of its 128-row block; a row block is complete when its
   * counter reaches ceil(N / tile width) (vqa_linear_tiles_n).  Lets a consumer kernel on another stream start on
   * finished row blocks while the GEMM is still running (vqa_graph_attention_args.d_progress). */
  int* d_progress;
  /* dtype / out_dtype VQA_F16X2: byte offset of the lo' plane behind the hi plane that d_A / d_W / d_out point to;
   * 0 = the planes are adjacent (M·lda·2, N·ldw·2, M·ldo·2 bytes) */
  size_t a_plane, w_plane, out_plane;
} vqa_linear_args;

int vqa_linear(const vqa_linear_args* args, void* stream);
int vqa_linear_part_width(int dtype);
/* number of output tiles (the unit of tile_begin / tile_end) the call would walk over; 0 when the shape takes a path
 * without tile ranges (fp32) */
int vqa_linear_tile_count(const vqa_linear_args* args);
/* tiles per 128-row block along N (the value a d_progress counter reaches when its row block is complete) */
int vqa_linear_tiles_n(const vqa_linear_args* args);
size_t vqa_linear_argmax_workspace_bytes(int M);

/* ------------------------------------------------------------------------
 * a6/a7  question encoder: embedding gather + 1-layer GRU, last state
 * replaces encoder.py:159-160, modules.py:139-159 (nn.Embedding + nn.GRU,
 * h0 = 0, output[:, -1], gate order [r;z;n]).
 *   tokens int64 [B,T]; emb [ntoken+1, ld_emb] (dtype); w_ih [3H, ld_emb]
 *   (zero padded to ld_emb), w_hh [3H,H] (dtype); biases f32 [3H].
 *   workspace: vqa_gru_workspace_bytes(B,T,H,ld_emb,dtype) bytes; whatever the caller passes BEYOND that many
 *   bytes is zero-filled by the call (vqa_forward parks the must-be-zero scratch of its fused answer selection there).
 *   out: h_last f32 [B,H]; if d_h_last_lp != NULL also written in `dtype`.
 * Sequence form (config 5, replaces modules.py:147-152 SentenceEmbedding.forward_all):
 *   d_x != NULL      : dense inputs [B*T, E_pad] (dtype) instead of tokens + embedding gather
 *   d_out_all != NULL: every hidden state, [B,T,H] in `dtype` (d_h_last may then be NULL;
 *                      d_h_last_lp is ignored).
 * ---------------------------------------------------------------------- */
typedef struct {
  const int64_t* d_tokens;
  int B, T, H, E_pad, ntoken_rows;
  int dtype;
  const void* d_emb;
  const void* d_w_ih;  const float* d_b_ih;
  const void* d_w_hh;  const float* d_b_hh;
  /* optional gate-interleaved copies (bf16, H % 64 == 0) that select the persistent fused
   * tcgen05 kernel (one launch for all T steps); NULL = generic per-step path.
   *   d_wx_packed [3H,E_pad], d_wh_packed [3H,H]: 192-row block j = rows of units
   *   [64j,64j+64) in gate order [r(64) | z(64) | n(64)];
   *   d_bias_packed f32 [4H] = b_ir+b_hr | b_iz+b_hz | b_in | b_hn            */
  const void* d_wx_packed; const void* d_wh_packed; const float* d_bias_packed;
  void* d_workspace;   size_t workspace_bytes;
  float* d_h_last;     void* d_h_last_lp;
  const void* d_x;     void* d_out_all;
  /* optional token table: row v = W_ih·emb[v] + bias, the input half of the gates, which depends on the token only
   * (modules.py:153 evaluates it per (sample, step)).
   *   dtype VQA_BF16: fp16 [ntoken_rows, H/32, 3, 32] — per 32-unit block the gates r | z | n, biases b_ir+b_hr |
   *     b_iz+b_hz | b_in folded in.  With the packed weights and tokens (no d_x / d_out_all) it selects the token-table
   *     form of the fused kernel: no embedding gather, no x-part GEMM — the kernel stages row tokens[b,t] of the table
   *     in shared memory (one bulk copy per row and step) and starts the accumulators from it.  NULL = off.
   *   dtype VQA_F16X2: f32 [ntoken_rows, 3H] in torch's gate order, b_ih folded in; required (this mode has no x-part). */
  const void* d_gi_table;
} vqa_gru_args;

int vqa_gru_last_state(const vqa_gru_args* args, void* stream);
size_t vqa_gru_workspace_bytes(int B, int T, int H, int E_pad, int dtype);

/* ------------------------------------------------------------------------
 * a7 (rnn_type='LSTM')  1-layer LSTM over dense inputs, h0 = c0 = 0
 * replaces modules.py:121-130,139-159 (nn.LSTM, gate order [i;f;g;o]; SentenceEmbedding.forward_all /
 * forward = output[:, -1]).  Per-step path: one GEMM for the input half of all steps, then per step one
 * GEMM (recurrent half + the input half as its additive epilogue operand) and the gate kernel.
 *   x [B*T, E_pad] (dtype, zero padded); w_ih [4H, E_pad], w_hh [4H, H] (dtype); biases f32 [4H]
 *   outputs (at least one): out_all [B,T,H] (dtype) every hidden state; h_last f32 [B,H]
 *   workspace: vqa_lstm_workspace_bytes(B,T,H,dtype) bytes.
 * ---------------------------------------------------------------------- */
typedef struct {
  int B, T, H, E_pad, dtype;
  const void* d_x;
  const void* d_w_ih;  const float* d_b_ih;
  const void* d_w_hh;  const float* d_b_hh;
  void* d_workspace;   size_t workspace_bytes;
  float* d_h_last;     void* d_out_all;
} vqa_lstm_args;
int vqa_lstm_sequence(const vqa_lstm_args* args, void* stream);
size_t vqa_lstm_workspace_bytes(int B, int T, int H, int dtype);

/* ------------------------------------------------------------------------
 * k6  top-down attention pooling
 * replaces attention.py:86 (softmax over the K regions), encoder.py:166
 * (v = v_att * v) and predictor.py:85 (v.sum(1)).
 *   logit_parts f32 [B*K, n_parts] (from vqa_linear's reduction form),
 *   logit_bias = attention.linear.bias; x [B,K,V] (dtype).
 *   outputs (each optional, NULL = skip):
 *     att   f32 [B,K]            softmax weights
 *     vsum  [B,V] (dtype)        sum_k att_k x_k
 *     vatt  [B,K,V] (dtype)      att_k x_k   (encoder output 'v')
 * V % 8 == 0, K <= 64.
 * ---------------------------------------------------------------------- */
int vqa_attention_pool(const float* d_logit_parts, int n_parts, float logit_bias,
                       const void* d_x, int B, int K, int V, int dtype,
                       float* d_att, void* d_vsum, void* d_vatt, void* stream);

/* ------------------------------------------------------------------------
 * k9+k10  relation-masked graph attention (one CorrelatedGraphConv layer + the
 * GCN's ReLU), after the wide projection of the RAW region features x.
 * replaces gcn.py:93-107 (conv, label bias), gcn.py:119-128 (relation_alpha),
 * gcn.py:152-168 (forward), gcn.py:211-212 (dropout=identity, ReLU) and
 * predictor.py:85 (the K-sum) when vsum is requested.
 *
 * layout 0 (f32 FFMA kernel, also bf16):  Y = x * [W0+W1 ; W2 ; Wa ; Wb]^T
 *   Y [B*K, ldy] (dtype): columns [0,V) P=(W0+W1)x, [V,2V) S=W2 x,
 *                         [2V,3V) Wa x, [3V,4V) Wb x ; ba, bb f32 [V]
 * layout 1 (bf16 only, tcgen05 kernel):   Y = x * [W0+W1 ; W2 ; Wb^T Wa]^T
 *   Y [B*K, ldy] bf16: [0,V) P, [V,2V) S, [2V,3V) Q with Q_i . x_j = (Wa x_i).(Wb x_j)
 *   x [B*K, ldx] bf16 (the raw features again: second operand of Q x^T)
 *   wvec bf16 [16,V]: row 0 = Wa^T bb, row 1 = Wb^T ba, rows 2..15 zero
 *   c0 = ba . bb ; label_bias_lp bf16 [16,V] (rows >= num_labels zero)
 *   K == 36, V % 128 == 0.
 * common:
 *   att f32 [B,K]: the top-down attention (feature f_i = att_i * x_i; row
 *                  scaling commutes with the bias-free maps), NULL = all ones
 *   labels u8 [B,K,K]; label_bias f32 [L,V] (layout 0)
 *   outputs (optional): out [B,K,V] (dtype) = ReLU(alpha . conv);
 *                       vsum [B,V] (dtype) = sum_i out_i; alpha f32 [B,K,K]
 * ---------------------------------------------------------------------- */
typedef struct {
  const void* d_Y; int ldy;
  const float* d_att;
  const uint8_t* d_labels;
  const float* d_label_bias; int num_labels;
  const float* d_ba; const float* d_bb;
  int B, K, V, dtype;
  void* d_out; void* d_vsum; float* d_alpha;
  /* layout 1 (merged algebra) */
  int layout;
  const void* d_x; int ldx;
  const void* d_wvec;
  float c0;
  const void* d_label_bias_lp;
  /* layout 1 only, optional: the kernel runs BESIDE the GEMM that produces Y (another stream, grid capped at cta_limit
   * SMs) and starts on an image as soon as the 128-row blocks of Y that hold its K rows are complete: d_progress are
   * that GEMM's counters (vqa_linear_args.d_progress), progress_target = vqa_linear_tiles_n of it.  Images are taken
   * in ascending order, the order in which the GEMM finishes them, so Y is read while it is still in L2.
   * The producer GEMM must be running (or done): a wait of more than ~2 s traps instead of hanging. */
  const int* d_progress; int progress_target;
  int cta_limit;             /* 0 = one CTA per SM */
} vqa_graph_attention_args;

int vqa_graph_attention(const vqa_graph_attention_args* args, void* stream);
/* dst += src, n elements of `dtype` (n % 8 == 0, 16-byte aligned): the sum of the implicit and the spatial relation
 * branch, encoder.py:257,264 (RelationEncoder with use_imp=True) */
int vqa_add_inplace(void* d_dst, const void* d_src, size_t n, int dtype, void* stream);

/* ------------------------------------------------------------------------
 * a14  answer selection: lowest-index argmax over the A logits
 * replaces wrapper.py:14 (torch.max(predict, 1)[1]).
 * ---------------------------------------------------------------------- */
int vqa_argmax_rows(const float* d_logits, int B, int A, int ld, int64_t* d_label,
                    void* stream);
/* VQA soft score of the chosen answers, replaces wrapper.py:16-22 (one_hot(label) * target) and the
 * `.sum()` the evaluation loop takes of it (train.py:186-189), with no host round trip:
 *   d_scores_dense [B,A] f32 or NULL;  d_score_row [B] f32 = target[b, label[b]] or NULL;
 *   d_score_sum [1] f32 = sum_b score_row[b] (fixed summation order) or NULL (needs d_score_row). */
int vqa_answer_scores(const int64_t* d_label, const float* d_target, int B, int A, int ld_target,
                      float* d_scores_dense, float* d_score_row, float* d_score_sum, void* stream);

/* ------------------------------------------------------------------------
 * config 5 (predictor_type 'q-cap') glue between the GEMMs and the two caption GRUs
 * (all tensors contiguous, H % 8 == 0, 16-byte aligned; `dtype` = element type of the
 * non-f32 tensors):
 *  vqa_caption_gate_scale  replaces modules.py:225-243 (CaptionAttention) + :294-295:
 *      a = sigmoid(h_w*p + h_w*r), h_w = out_w[:, T-1, :];  in2[b,t,:] = a[b,:] * out_w[b,t,:]
 *      out_w, in2 [B,T,H] (dtype); p, r f32 [B,H] = LReLU(W_v v), LReLU(W_q q); d_a f32 [B,H] optional
 *  vqa_seq_max             replaces modules.py:306 (output.max(dim=1)[0]): [B,T,H] -> [B,H] (dtype)
 *  vqa_softmax_mul         replaces predictor.py:202-203: out = softmax_H(z) * v, z f32 [B,H],
 *      v, out [B,H] (dtype)   (sum_K (joint*V_k) == joint * sum_K V_k, so v is the pooled feature)
 * ---------------------------------------------------------------------- */
int vqa_caption_gate_scale(const void* d_out_w, const float* d_p, const float* d_r, int B, int T, int H,
                           int dtype, void* d_in2, float* d_a, void* stream);
int vqa_seq_max(const void* d_e, int B, int T, int H, int dtype, void* d_out, void* stream);
int vqa_softmax_mul(const float* d_z, const void* d_v, int B, int H, int dtype, void* d_out, void* stream);

/* ------------------------------------------------------------------------
 * caption head (decoder_type 'base'): per-step pieces of BaseDecoder.decode
 * (generator.py:168-181); the GEMMs are vqa_linear, the softmax + weighted sum is
 * vqa_attention_pool.
 *  vqa_attention_logits  replaces attention.py:70-76 / :33-40 evaluated at every decoding step: the region
 *      half of the attention is projected ONCE per caption batch (it does not depend on the hidden state),
 *      each step reduces it against the hidden-state half.
 *      proj [B*K, ldp] (dtype), q f32 [B, ldq], w f32 [Hd] (the weight-normed 1-wide logit layer),
 *      logits f32 [B*K] (bias not added: it is vqa_attention_pool's logit_bias; n_parts = 1)
 *        mode 0 ('new'):  proj = ReLU(W_v v + b), q = ReLU(W_q h + b):  sum_h proj*q*w
 *        mode 1 ('base'): proj = W1[:, :V] v,     q = W1[:, V:] h + b1: sum_h ReLU(proj + q)*w
 *      Hd, ldp, ldq % 8 == 0, 16-byte aligned bases.
 *  vqa_gru_cell          replaces nn.GRUCell's gate update (generator.py:158-159,178), gate order [r;z;n]:
 *      gi = W_ih x + b_ih, gh = W_hh h + b_hh  f32 [B,3H] (from vqa_linear);  h' = (1-z)*n + z*h
 *      h_prev, h_out f32 [B,H] (may alias); h_lp [B, ld_lp] (dtype): the copy the next GEMMs read.
 * ---------------------------------------------------------------------- */
int vqa_attention_logits(const void* d_proj, int ldp, const float* d_q, int ldq, const float* d_w, int B, int K,
                         int Hd, int mode, int dtype, float* d_logits, void* stream);
int vqa_gru_cell(const float* d_gi, const float* d_gh, const float* d_h_prev, int B, int H, int dtype,
                 float* d_h_out, void* d_h_lp, int ld_lp, void* stream);
/*  vqa_lstm_cell         nn.LSTMCell's gate update: gates f32 [B,4H] = W_ih x + b_ih + W_hh h + b_hh ([i;f;g;o]);
 *      c f32 [B,H] updated in place; h_out f32 [B,H] (optional); h_lp [B, ld_lp] (dtype). */
int vqa_lstm_cell(const float* d_gates, int B, int H, int dtype, float* d_c, float* d_h_out, void* d_h_lp, int ld_lp,
                  void* stream);

/*  vqa_caption_decode_steps  replaces the time loop of DecoderModule.forward (generator.py:99-111) around
 *      BaseDecoder.decode (:168-181) in ONE call: T teacher-forced steps, step t on the first h_batches[t]
 *      samples (captions sorted by decreasing length, so h_batches is non-increasing).  Per step:
 *      q = act(W_q h) -> vqa_attention_logits -> vqa_attention_pool -> gi = W_att att_v + gi_prev[:, t]
 *      -> gh = W_hh h + b_hh -> GRUCell gate update.  Hoisted by the caller (they do not depend on h):
 *      proj (region half of the attention), gi_prev (previous-word half of the GRUCell input GEMM, all steps);
 *      the word logits Linear(h_t) of all steps are one vqa_linear over h_all afterwards.
 *        x [B,K,V], proj [B*K,Hd], w_q [Hd,Hd] (+ scale/bias f32 [Hd]), w_att [3Hd,V], w_hh [3Hd,Hd] (dtype);
 *        gi_prev f32 [B, T*3Hd]; h f32 [B,Hd] in: initial state, out: final states; h0_lp [B,Hd] (dtype) = h;
 *        h_all out [sum_t h_batches[t], Hd] (dtype): every new state, time-major = pack_padded_sequence order.
 *      h_batches is a HOST array.  Hd % 8 == 0, V % 8 == 0. */
typedef struct {
  int B, K, V, Hd, T, dtype;
  const int* h_batches;
  const void* d_x; const void* d_proj;
  int att_mode;                       /* 0 = MultiplyAttention, 1 = ConcatAttention (see vqa_attention_logits) */
  const void* d_wq; const float* d_wq_scale; const float* d_wq_bias;
  const float* d_logit_w; float logit_bias;
  const float* d_gi_prev;
  const void* d_w_att; const void* d_w_hh; const float* d_b_hh;
  void* d_h_all; float* d_h; const void* d_h0_lp;
  void* d_workspace; size_t workspace_bytes;
  /* rnn_type: 0 = nn.GRUCell (gate blocks of 3Hd: gi_prev [B, T*3Hd], w_att [3Hd,V], w_hh [3Hd,Hd]);
   *           1 = nn.LSTMCell (4Hd everywhere, gate order [i;f;g;o]); d_c f32 [B,Hd]: cell state, in/out */
  int cell; float* d_c;
} vqa_caption_decode_args;
size_t vqa_caption_decode_workspace_bytes(int B, int K, int V, int Hd, int dtype);
int vqa_caption_decode_steps(const vqa_caption_decode_args* args, void* stream);

/* ------------------------------------------------------------------------
 * optimizer side of the training step (BASELINE config 4; the step after the path):
 * one launch per operation over ALL parameter tensors (pointer table by value).
 *  vqa_grad_clip     replaces nn.utils.clip_grad_norm_(model.parameters(), max_norm) (train.py:109):
 *      total_norm = ||all grads||_2, scale = min(1, max_norm / (total_norm + 1e-6)), both written to DEVICE
 *      scalars (no host sync); scale_in_place != 0 also multiplies every gradient by scale (what the
 *      reference leaves in .grad), 0 leaves the gradients alone (hand d_scale to vqa_adamax_step instead).
 *      workspace: vqa_grad_clip_workspace_bytes() bytes.
 *  vqa_adamax_step   replaces torch.optim.Adamax(...).step() (train.py:57,110), maximize=False:
 *      g = grad * (*d_grad_scale) + weight_decay * p;  m += (1-beta1)(g-m);  u = max(beta2*u, |g|+eps);
 *      p -= lr / (1 - beta1^step) * m / u.     lr is per tensor (the param groups of train.py:54-56).
 *      All tensors f32, contiguous; d_m / d_u are the optimizer state (exp_avg / exp_inf).
 * h_tensors is a HOST array; at most 8 * VQA_OPTIM_MAX_TENSORS entries.
 * ---------------------------------------------------------------------- */
#define VQA_OPTIM_MAX_TENSORS 64
typedef struct {
  float* d_p;          /* parameter (NULL for vqa_grad_clip)  */
  const float* d_g;    /* gradient                            */
  float* d_m;          /* exp_avg  (NULL for vqa_grad_clip)   */
  float* d_u;          /* exp_inf  (NULL for vqa_grad_clip)   */
  size_t n;            /* elements                            */
  float lr;            /* learning rate of this tensor's group */
} vqa_optim_tensor;
size_t vqa_grad_clip_workspace_bytes(void);
int vqa_grad_clip(const vqa_optim_tensor* h_tensors, int n_tensors, float max_norm, int scale_in_place,
                  void* d_workspace, float* d_total_norm, float* d_scale, void* stream);
int vqa_adamax_step(const vqa_optim_tensor* h_tensors, int n_tensors, float beta1, float beta2, float eps,
                    float weight_decay, int step, const float* d_grad_scale, void* stream);

/* ------------------------------------------------------------------------
 * whole path: Wrapper.forward / forward_vqa (wrapper.py:64-74,113-118) for
 * encoder_type in {base, relation}, att_type in {'new', 'base'}, predictor 'base'.
 * All weight pointers are "prepared" device tensors (see
 * vqa_collection_b200/engine.py: bf16 or f32 copies, weight-norm scalars
 * expanded to per-column scale vectors, concatenated where noted).
 * ---------------------------------------------------------------------- */
typedef struct {
  /* dims */
  int B, K, V, H, A, T, E_pad, ntoken_rows, num_labels;
  int dtype;                 /* vqa_dtype of activations and weights         */
  int relation;              /* 0 = Up-Down, 1 = Up-Down + ReGAT (1 layer)   */
  /* inputs */
  const void* d_img;         /* [B,K,V] dtype                                */
  const int64_t* d_tokens;   /* [B,T]                                        */
  const uint8_t* d_labels;   /* [B,K,K] (relation) or NULL                   */
  const float* d_bbox;       /* [B,K,4]: if non-NULL and d_labels_out given,
                                labels are computed on device first          */
  float img_w, img_h;
  /* question encoder */
  const void* d_emb; const void* d_w_ih; const float* d_b_ih;
  const void* d_w_hh; const float* d_b_hh;
  const void* d_wx_packed; const void* d_wh_packed; const float* d_bias_packed;  /* see vqa_gru_args */
  /* attention (attention.py:61-66) */
  const void* d_Wv;  const float* d_sv;  const float* d_bv;      /* [H,V]    */
  const void* d_Wqq; const float* d_sqq; const float* d_bqq;     /* [2H,H]: W_q ; q_net */
  const float* d_wlin;       /* [H] = linear.weight_v * g/||v||              */
  float b_lin;
  /* att_type 'base' (ConcatAttention, attention.py:18-51): att_concat = 1, then
   *   d_Wv/d_sv = the v-half W1[:, :V] of sequence.0 and its scale (d_bv unused),
   *   d_W1q [H,H] = the q-half W1[:, V:], d_b1 = sequence.0.bias (scale = d_sv),
   *   d_Wqq/d_sqq/d_bqq = q_net alone [H,H], d_wlin/b_lin = sequence.2           */
  int att_concat; const void* d_W1q; const float* d_b1;
  /* ReGAT layer (gcn.py), concatenated [4V,V] = [W0+W1; W2; Wa; Wb] */
  const void* d_Wg; const float* d_label_bias; const float* d_ba; const float* d_bb;
  /* bf16 merged form (vqa_graph_attention layout 1): [3V,V] = [W0+W1; W2; Wb^T Wa]; when
   * d_Wg3 is non-NULL (bf16 only) it replaces d_Wg/d_ba/d_bb/d_label_bias */
  const void* d_Wg3; const void* d_wvec; float gat_c0; const void* d_label_bias_lp;
  /* predictor (predictor.py:58-79) */
  const void* d_Wvn; const float* d_svn; const float* d_bvn;     /* [H,V]    */
  const void* d_Wc0; const float* d_sc0; const float* d_bc0;     /* [2H,H]   */
  const void* d_Wc1; const float* d_sc1; const float* d_bc1;     /* [A,2H], rows padded to mult of 8 */
  /* workspace */
  void* d_workspace; size_t workspace_bytes;
  /* outputs (optional ones may be NULL) */
  float* d_logits;           /* [B,A] f32                                    */
  int64_t* d_label;          /* [B]   lowest-index argmax                    */
  float* d_att;              /* [B,K] attention weights (v_att)              */
  float* d_q;                /* [B,H] q_net output, optional                 */
  void* d_v;                 /* [B,K,V] encoder output 'v', optional         */
  float* d_alpha;            /* [B,K,K], optional (relation)                 */
  uint8_t* d_labels_out;     /* [B,K,K], optional (relation + bbox)          */
  /* Two-stream schedule (bf16 tensor-core path, B >= 512; 0 = every kernel on `stream`, in order).
   * overlap = 1: the question encoder (gather, GRU on side_sms SMs with two interleaved row blocks per CTA pair,
   * [W_q ; q_net]) runs on a library-owned side stream WHILE the question-independent projection of the region
   * features runs on `stream` on the other SMs — ReGAT: x·[W0+W1; W2; WbᵀWa]ᵀ; Up-Down: ReLU(W_v x) stored as
   * bf16 [B·K,H] (the ⊙q logit reduction then is its own streaming kernel).  When the encoder is done, its SMs
   * compute the last side_tile_permille / 1000 of the projection's tiles.  The streams are joined before the
   * attention; the call can be captured in a CUDA graph (the first, uncaptured call creates the side stream). */
  int overlap;
  int side_sms;              /* SMs of the side stream, even (0 = 64)                                        */
  int side_tile_permille;    /* 0 = automatic (from the shapes); -1 = none                                   */
  /* ReGAT, bf16 merged form: gat_chase_sms = G > 0 runs the graph attention on G SMs of the side stream BESIDE the wide
   * projection (which keeps the other SMs): the GEMM publishes every finished 128-row block of Y (vqa_linear_args
   * .d_progress) and the graph attention works on an image as soon as its rows are complete, i.e. while they are still in
   * L2 — Y's HBM read and most of the HBM-bound kernel's SM time disappear behind the tensor-bound one.  Takes
   * precedence over `overlap`.  Falls back to the serial order under a profiler / CUDA_LAUNCH_BLOCKING (kernels of the
   * two streams must be able to run at the same time). */
  int gat_chase_sms;
  /* question encoder, token-table form: see vqa_gru_args.d_gi_table (NULL = gather + x-part GEMM) */
  const void* d_gi_table;
} vqa_forward_args;

size_t vqa_forward_workspace_bytes(const vqa_forward_args* args);
int vqa_forward(const vqa_forward_args* args, void* stream);
/* number of kernels the last vqa_forward on this thread launched */
int vqa_forward_last_launch_count(void);

/* ------------------------------------------------------------------------
 * whole path from HOST buffers in the reference's wire format (the e2e leg).
 * replaces the `.to(self.device)` copies inside the reference's forwards
 * (encoder.py:153-156,265; predictor.py:82-83) + Wrapper.forward_vqa.
 *   h_img    f32 [B,K,V] (dataset.py:96-104; pinned memory recommended)
 *   h_tokens int64 [B,T];  h_labels u8 [B,K,K] or h_bbox f32 [B,K,4] (relation)
 *   h_label  int64 [B] out (the answers); synchronises `stream` before returning
 * `fwd` carries dims, weights, workspace and optional DEVICE outputs exactly as
 * for vqa_forward; its d_img / d_tokens / d_labels / d_bbox / d_label members are
 * ignored (the context owns those buffers).
 * pack_on_host (bf16 engines): 1 = the host cores convert f32 -> bf16 (round to
 * nearest even, bit-identical to vqa_cast_f32_to_bf16) into pinned staging slots
 * while the previous chunk is in flight, so PCIe carries 2 bytes per feature;
 * 0 = f32 over PCIe + device cast.  chunk_rows = images per staged chunk (0 = 64).
 * raw_chunk_period = n mixes the two so that neither the host memory system (packing)
 * nor PCIe (raw f32) is the lone bottleneck.
 * h2d_bytes / d2h_bytes report what crossed PCIe.
 * ---------------------------------------------------------------------- */
typedef struct vqa_host_ctx vqa_host_ctx;
int vqa_host_ctx_create(vqa_host_ctx** ctx, int pack_threads /* 0 = all host cores */);
void vqa_host_ctx_destroy(vqa_host_ctx* ctx);
int vqa_host_ctx_threads(vqa_host_ctx* ctx);

typedef struct {
  vqa_forward_args fwd;
  const float* h_img;
  const int64_t* h_tokens;
  const uint8_t* h_labels;
  const float* h_bbox;
  int64_t* h_label;
  int chunk_rows;
  int pack_on_host;
  int raw_chunk_period;      /* with pack_on_host: every n-th chunk (n >= 2) goes as f32 + device cast, 0 = none */
  size_t h2d_bytes, d2h_bytes;
  int img_is_bf16;           /* h_img holds bf16 [B,K,V] (a feature cache kept in the resident format):
                                chunks go straight into HBM, no host pack and no device cast (dtype must be VQA_BF16) */
} vqa_forward_host_args;

int vqa_forward_host(vqa_host_ctx* ctx, vqa_forward_host_args* args, void* stream);
/* split form for pipelining batches over two contexts: submit stages + enqueues everything and returns
 * without synchronising; wait blocks until the answers of the submitted batch are in h_label.  While batch
 * n computes, the host cores can already pack batch n+1 of the other context.                           */
int vqa_forward_host_submit(vqa_host_ctx* ctx, vqa_forward_host_args* args, void* stream);
int vqa_forward_host_wait(vqa_host_ctx* ctx);

/* ------------------------------------------------------------------------
 * training step of the Up-Down path (BASELINE config 4): forward with saved
 * activations + BCE loss + full backward in one call.
 * replaces Wrapper.get_loss (wrapper.py:76-105: forward, instance_bce_with_logits
 * :25-29) followed by loss.backward() (train.py:108) for encoder 'base',
 * att_type 'new', predictor 'base'.  The caller (train.py:109-111, unchanged)
 * still owns gradient clipping and the Adamax step; with several GPUs the
 * gradients are all-reduced between this call and the clip
 * (vqa_collection_b200/parallel.py).
 *
 * Parameters are the f32 MASTER tensors under the reference's names and layouts;
 * gradients are written (not accumulated) as f32 tensors of the same shapes.
 * Layer order of p_v/p_g/p_b/g_v/g_g/g_b (weight_v, weight_g, bias):
 *   0 encoder.attention.W_v.main.0 [H,V]   1 encoder.attention.W_q.main.0 [H,H]
 *   2 encoder.attention.linear [1,H]       3 encoder.q_net.main.0 [H,H]
 *   4 predictor.v_net.main.0 [H,V]         5 predictor.classifier.main.0 [2H,H]
 *   6 predictor.classifier.main.3 [A,2H]
 * dropout_att (attention.py:74) and dropout_cls (modules.py:45) are the train-
 * mode dropout probabilities; masks come from a counter-based hash of `seed`.
 * loss = mean(BCEWithLogits(logits, target)) * A, logits = the ReLU'd answers.
 * ---------------------------------------------------------------------- */
#define VQA_TRAIN_LAYERS 7
typedef struct {
  int B, K, V, H, A, T, E, ntoken_rows;
  int dtype;                          /* compute dtype of operands/activations */
  float dropout_att, dropout_cls;
  unsigned long long seed;
  const void* d_img;                  /* [B,K,V] dtype                         */
  const int64_t* d_tokens;            /* [B,T]                                 */
  const float* d_target;              /* [B,A] soft scores                     */
  const float* p_emb;                 /* [ntoken_rows,E]; last row = padding   */
  const float* p_w_ih; const float* p_w_hh; const float* p_b_ih; const float* p_b_hh;
  const float* p_v[VQA_TRAIN_LAYERS]; const float* p_g[VQA_TRAIN_LAYERS]; const float* p_b[VQA_TRAIN_LAYERS];
  float* g_emb; float* g_w_ih; float* g_w_hh; float* g_b_ih; float* g_b_hh;
  float* g_v[VQA_TRAIN_LAYERS]; float* g_g[VQA_TRAIN_LAYERS]; float* g_b[VQA_TRAIN_LAYERS];
  float* d_loss;                      /* [1]                                   */
  float* d_logits;                    /* [B,A] f32 predictions                 */
  void* d_workspace; size_t workspace_bytes;
  /* optional cudaEvent_t, recorded on `stream` when the gradients of the seven weight-normed layers (g_v / g_g / g_b:
   * 59 of the 75.5 MB) are final — before the back-propagation through time — so that a data-parallel caller can
   * all-reduce that bucket on another stream while the GRU / embedding gradients are still being computed */
  void* ev_head_done;
} vqa_train_args;

size_t vqa_train_workspace_bytes(const vqa_train_args* args);
int vqa_updown_train_step(const vqa_train_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VQA_B200_H_ */
