// k6 top-down attention pooling, a14 lowest-index argmax, GRU gate update,
// embedding gather and f32<->bf16 casts.  All HBM-bound streaming kernels:
// 16-byte vector accesses, one CTA per image / row group, no tensor cores.
#include "common.cuh"

namespace vqa {

// ---------------------------------------------------------------------------
// k6: att = softmax_K(sum_p parts + b); vsum = sum_k att_k x_k; vatt = att_k x_k
// Reference: attention.py:86, encoder.py:166, predictor.py:85.
// ---------------------------------------------------------------------------
constexpr int kPoolThreads = 64;           // 2 warps; one CTA = (image, 512-channel slice)
constexpr int kPoolChan = kPoolThreads * 8;
constexpr int kPoolMaxK = 64;

// Work item = (image, 512-channel slice): B·V/512 small CTAs (4096 at B=1024) so that the
// whole grid is resident in one wave (no wave tail) and every SM keeps ~1600 threads with
// four 16-byte loads each in flight.  Every CTA recomputes the 36-way softmax (L2 hits).
template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
attention_pool_kernel(const float* __restrict__ parts, int n_parts, float bias,
                      const T* __restrict__ x, int B, int K, int V, int slices, int rev,
                      float* __restrict__ att_out, T* __restrict__ vsum, T* __restrict__ vatt) {
  __shared__ float s_att[kPoolMaxK];
  griddep_launch();
  griddep_wait();
  // images in DESCENDING order: the projection GEMM that ran just before swept the same features in ascending row
  // order, so its last ~100 MB are still in the 126 MB L2 when the first CTAs of this kernel ask for them
  const int b = rev ? B - 1 - (int)(blockIdx.x / slices) : (int)(blockIdx.x / slices);
  const int slice = blockIdx.x % slices;
  const int tid = threadIdx.x;
  if (tid < 32) {
    // K <= 64: lane handles k = lane and k = lane + 32
    float l0 = -INFINITY, l1 = -INFINITY;
    if (tid < K) {
      const float* p = parts + (size_t)(b * K + tid) * n_parts;
      float s = 0.f;
      for (int i = 0; i < n_parts; ++i) s += __ldg(p + i);
      l0 = s + bias;
    }
    if (tid + 32 < K) {
      const float* p = parts + (size_t)(b * K + tid + 32) * n_parts;
      float s = 0.f;
      for (int i = 0; i < n_parts; ++i) s += __ldg(p + i);
      l1 = s + bias;
    }
    const float m = warp_max(fmaxf(l0, l1));
    const float e0 = (tid < K) ? expf(l0 - m) : 0.f;
    const float e1 = (tid + 32 < K) ? expf(l1 - m) : 0.f;
    const float inv = 1.f / warp_sum(e0 + e1);
    if (tid < K) s_att[tid] = e0 * inv;
    if (tid + 32 < K) s_att[tid + 32] = e1 * inv;
  }
  __syncthreads();
  if (att_out != nullptr && slice == 0 && tid < K) att_out[(size_t)b * K + tid] = s_att[tid];
  if (vsum == nullptr && vatt == nullptr) return;

  const int c = slice * kPoolChan + tid * 8;
  if (c >= V) return;
  const T* xb = x + (size_t)b * K * V + c;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 6
  for (int k = 0; k < K; ++k) {
    float v[8];
    load8(xb + (size_t)k * V, v);
    const float a = s_att[k];
    if (vatt != nullptr) {
      float w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = a * v[i];
      store8(vatt + ((size_t)b * K + k) * V + c, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += w[i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, v[i], acc[i]);
    }
  }
  if (vsum != nullptr) store8(vsum + (size_t)b * V + c, acc);
}

int attention_pool(const float* parts, int n_parts, float bias, const void* x, int B, int K, int V,
                   int dtype, float* att, void* vsum, void* vatt, cudaStream_t s) {
  VQA_REQUIRE(K >= 1 && K <= kPoolMaxK, "attention_pool: K=%d out of range", K);
  VQA_REQUIRE(V % 8 == 0 && n_parts >= 1, "attention_pool: V=%d must be a multiple of 8", V);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(parts && x, "attention_pool: NULL input");
  const bool stream_x = vsum != nullptr || vatt != nullptr;
  const int slices = stream_x ? (V + kPoolChan - 1) / kPoolChan : 1;
  const unsigned grid = (unsigned)B * slices;
  const int rev = l2_order_enabled() ? 1 : 0;
  if (dtype == VQA_BF16) {
    VQA_CUDA_CHECK(launch_pdl(attention_pool_kernel<__nv_bfloat16>, dim3(grid), dim3(kPoolThreads), 0, s,
                              parts, n_parts, bias, (const __nv_bfloat16*)x, B, K, V, slices, rev, att, (__nv_bfloat16*)vsum,
                              (__nv_bfloat16*)vatt));
  } else {
    VQA_CUDA_CHECK(launch_pdl(attention_pool_kernel<float>, dim3(grid), dim3(kPoolThreads), 0, s, parts, n_parts, bias,
                              (const float*)x, B, K, V, slices, rev, att, (float*)vsum, (float*)vatt));
  }
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// ---------------------------------------------------------------------------
// a14: lowest-index argmax per row (wrapper.py:14; torch.max tie rule).
// One warp per row.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
argmax_rows_kernel(const float* __restrict__ logits, int B, int A, int ld, int64_t* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  griddep_launch();
  griddep_wait();
  if (row >= B) return;
  const float* p = logits + (size_t)row * ld;
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int n = lane; n < A; n += 32) {
    const float v = p[n];
    if (v > best || idx == 0x7fffffff) { best = v; idx = n; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
  }
  if (lane == 0) out[row] = (int64_t)idx;
}

int argmax_rows(const float* logits, int B, int A, int ld, int64_t* out, cudaStream_t s) {
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(logits && out && A >= 1 && ld >= A, "argmax_rows: bad arguments");
  const int rows_per_cta = 8;
  VQA_CUDA_CHECK(launch_pdl(argmax_rows_kernel, dim3((B + rows_per_cta - 1) / rows_per_cta), dim3(rows_per_cta * 32), 0, s,
                            logits, B, A, ld, out));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// ---------------------------------------------------------------------------
// embedding gather (encoder.py:159): X[b*T+t, :] = emb[token, :], rows of E_pad
// elements copied as 16-byte vectors.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embedding_gather_kernel(const int64_t* __restrict__ tokens, int n_rows, int row_vec16, int ntoken_rows,
                        const uint4* __restrict__ emb, uint4* __restrict__ out) {
  const int total = n_rows * row_vec16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / row_vec16, c = i - r * row_vec16;
    long long tok = tokens[r];
    tok = tok < 0 ? 0 : (tok >= ntoken_rows ? ntoken_rows - 1 : tok);
    out[i] = __ldg(emb + (size_t)tok * row_vec16 + c);
  }
}

int embedding_gather(const int64_t* tokens, int n_rows, int E_pad, int ntoken_rows, int dtype,
                     const void* emb, void* out, cudaStream_t s) {
  const size_t row_bytes = (size_t)E_pad * elem_size(dtype);
  VQA_REQUIRE(row_bytes % 16 == 0, "embedding_gather: padded row of %zu bytes is not 16-byte aligned",
              row_bytes);
  if (n_rows == 0) return VQA_OK;
  const int row_vec16 = (int)(row_bytes / 16);
  const int total = n_rows * row_vec16;
  int grid = (total + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  embedding_gather_kernel<<<grid, 256, 0, s>>>(tokens, n_rows, row_vec16, ntoken_rows,
                                               (const uint4*)emb, (uint4*)out);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// ---------------------------------------------------------------------------
// GRU gate update (torch GRU semantics, modules.py:153):
//   r = σ(gi_r + gh_r), z = σ(gi_z + gh_z), n = tanh(gi_n + r ⊙ gh_n)
//   h' = (1 - z) ⊙ n + z ⊙ h          (gi, gh already contain b_ih, b_hh)
// gi row of sample b at step t is (b*T + t); h kept in f32, plus a low
// precision copy that is the next step's GEMM operand.
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
gru_gate_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int B, int H, int Tlen,
                int t, const float* h_prev, float* h_out, T* __restrict__ h_lp, int ld_lp) {
  const int total = B * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / H, j = i - b * H;
    const float* gir = gi + ((size_t)b * Tlen + t) * 3 * H;
    const float* ghr = gh + (size_t)b * 3 * H;
    const float r = 1.f / (1.f + expf(-(gir[j] + ghr[j])));
    const float z = 1.f / (1.f + expf(-(gir[H + j] + ghr[H + j])));
    const float n = tanhf(gir[2 * H + j] + r * ghr[2 * H + j]);
    const float hn = (1.f - z) * n + z * h_prev[i];
    h_out[i] = hn;
    h_lp[(size_t)b * ld_lp + j] = Elem<T>::from_f(hn);
  }
}

// h_lp row stride ld_lp: H for the [B,H] operand copy, T*H when the states go straight into [B,T,H]
int gru_gate(const float* gi, const float* gh, int B, int H, int T, int t, const float* h_prev,
             float* h_out, void* h_lp, int ld_lp, int dtype, cudaStream_t s) {
  const int total = B * H;
  int grid = (total + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  if (dtype == VQA_BF16)
    gru_gate_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(gi, gh, B, H, T, t, h_prev, h_out,
                                                        (__nv_bfloat16*)h_lp, ld_lp);
  else
    gru_gate_kernel<float><<<grid, 256, 0, s>>>(gi, gh, B, H, T, t, h_prev, h_out, (float*)h_lp, ld_lp);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// ---------------------------------------------------------------------------
// a14 / f4: VQA soft score of the chosen answers (wrapper.py:16-22: one_hot(label) ⊙ target) without the
// zeros → scatter → multiply round trips: one block per question writes its dense row (optional) and
// score_row[b] = target[b, label[b]]; a single warp then sums the rows in a fixed order (deterministic).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
answer_scores_kernel(const int64_t* __restrict__ label, const float* __restrict__ target, int A, int ld,
                     float* __restrict__ dense, float* __restrict__ score_row) {
  const int b = blockIdx.x;
  const long long l = label[b];
  const float v = (l >= 0 && l < A) ? target[(size_t)b * ld + l] : 0.f;
  if (dense != nullptr) {
    float* row = dense + (size_t)b * A;
    for (int n = threadIdx.x; n < A; n += 256) row[n] = (n == l) ? v : 0.f;
  }
  if (score_row != nullptr && threadIdx.x == 0) score_row[b] = v;
}
__global__ void __launch_bounds__(32) score_sum_kernel(const float* __restrict__ score_row, int B, float* __restrict__ out) {
  double acc = 0.0;
  for (int i = threadIdx.x; i < B; i += 32) acc += (double)score_row[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) out[0] = (float)acc;
}

int answer_scores(const int64_t* label, const float* target, int B, int A, int ld, float* dense, float* score_row,
                  float* score_sum, cudaStream_t s) {
  VQA_REQUIRE(B >= 0 && A >= 1 && ld >= A, "answer_scores: bad dims B=%d A=%d ld=%d", B, A, ld);
  VQA_REQUIRE(score_sum == nullptr || score_row != nullptr, "answer_scores: d_score_sum needs d_score_row");
  if (B == 0) {
    if (score_sum) VQA_CUDA_CHECK(cudaMemsetAsync(score_sum, 0, sizeof(float), s));
    return VQA_OK;
  }
  VQA_REQUIRE(label && target, "answer_scores: NULL input");
  answer_scores_kernel<<<B, 256, 0, s>>>(label, target, A, ld, dense, score_row);
  VQA_LAUNCH_CHECK();
  if (score_sum) {
    score_sum_kernel<<<1, 32, 0, s>>>(score_row, B, score_sum);
    VQA_LAUNCH_CHECK();
  }
  return VQA_OK;
}

// ---------------------------------------------------------------------------
// casts (wire format f32 -> resident bf16 and back)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
  if (blockIdx.x == 0)
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __float2bfloat16_rn(src[i]);
}

__global__ void __launch_bounds__(256)
cast_bf16_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, size_t n) {
  const size_t n8 = n / 8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8;
       i += (size_t)gridDim.x * blockDim.x) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
  if (blockIdx.x == 0)
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) dst[i] = __bfloat162float(src[i]);
}

int cast_f32_to_bf16(const float* src, void* dst, size_t n, cudaStream_t s) {
  VQA_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "cast: pointers must be 16-byte aligned");
  if (n == 0) return VQA_OK;
  size_t grid = (n / 8 + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
  cast_f32_bf16_kernel<<<(unsigned)grid, 256, 0, s>>>(src, (__nv_bfloat16*)dst, n);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

int cast_bf16_to_f32(const void* src, float* dst, size_t n, cudaStream_t s) {
  VQA_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0), "cast: pointers must be 16-byte aligned");
  if (n == 0) return VQA_OK;
  size_t grid = (n / 8 + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
  cast_bf16_f32_kernel<<<(unsigned)grid, 256, 0, s>>>((const __nv_bfloat16*)src, dst, n);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
