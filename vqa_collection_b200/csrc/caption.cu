// Config 5 (q-cap predictor) glue kernels — the element-wise / small-reduction steps between the
// GEMMs and the two GRUs of PredictorwithCaption (predictor.py:186-213) and CaptionEmbedding
// (modules.py:291-306).  All three are HBM/L2-streaming kernels over [B,T,H] or [B,H] tensors:
// 16-byte accesses, one pass, nothing staged.
//
//   caption_gate_scale : a = σ(h_w ⊙ p + h_w ⊙ r) with h_w = out_w[:, T-1, :]   (CaptionAttention,
//                        modules.py:225-243; p = LReLU(W_v v), r = LReLU(W_q q) come from vqa_linear)
//                        and in2[b,t,:] = a[b,:] ⊙ out_w[b,t,:]                 (modules.py:294-295)
//   seq_max            : out[b,:] = max_t e[b,t,:]                               (modules.py:306)
//   softmax_mul        : out[b,:] = softmax_H(z[b,:]) ⊙ v[b,:]                   (predictor.py:202-203;
//                        Σ_K (joint ⊙ V_k) = joint ⊙ Σ_K V_k, so only the pooled v is needed)
//
// Caption decoder (BaseDecoder.decode, generator.py:168-181) per-step glue:
//   attention_logits   : the decoder attends over the SAME regions at every step, so the region half of its
//                        attention (ReLU(W_v v), or W1_v v for att_type='base') is projected once per caption batch
//                        and each step only reduces it against that step's hidden-state half:
//                          mode 0: logit[b,k] = Σ_h proj[b,k,h] · q[b,h] · w[h]            (attention.py:70-76)
//                          mode 1: logit[b,k] = Σ_h ReLU(proj[b,k,h] + q[b,h]) · w[h]      (attention.py:33-40)
//                        one warp per (b,k) row, 16-byte loads; HBM/L2-streaming over proj.
//   (the GRUCell gate update is gru_gate of pool.cu with T = 1)
#include "common.cuh"

namespace vqa {

template <typename T>
__global__ void __launch_bounds__(256)
attention_logits_kernel(const T* __restrict__ proj, int ldp, const float* __restrict__ q, int ldq,
                        const float* __restrict__ w, int rows, int K, int Hd, int mode, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const T* pr = proj + (size_t)row * ldp;
    const float* qr = q + (size_t)(row / K) * ldq;
    float acc = 0.f;
    for (int h0 = lane * 8; h0 < Hd; h0 += 256) {
      float pv[8], qv[8], wv[8];
      load8(pr + h0, pv);
      load8(qr + h0, qv);
      load8(w + h0, wv);
      if (mode == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += pv[j] * qv[j] * wv[j];
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += fmaxf(pv[j] + qv[j], 0.f) * wv[j];
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[row] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
caption_gate_scale_kernel(const T* __restrict__ out_w, const float* __restrict__ p, const float* __restrict__ r,
                          int B, int Tn, int H, T* __restrict__ in2, float* __restrict__ a_out) {
  const int H8 = H / 8;
  const size_t total = (size_t)B * H8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H8), h0 = (int)(i - (size_t)b * H8) * 8;
    const T* row = out_w + (size_t)b * Tn * H + h0;
    float hw[8], pv[8], rv[8], a[8];
    load8(row + (size_t)(Tn - 1) * H, hw);
    load8(p + (size_t)b * H + h0, pv);
    load8(r + (size_t)b * H + h0, rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.f / (1.f + expf(-(hw[j] * pv[j] + hw[j] * rv[j])));
    if (a_out) store8(a_out + (size_t)b * H + h0, a);
    for (int t = 0; t < Tn; ++t) {
      float x[8];
      load8(row + (size_t)t * H, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] *= a[j];
      store8(in2 + (size_t)b * Tn * H + (size_t)t * H + h0, x);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
seq_max_kernel(const T* __restrict__ e, int B, int Tn, int H, T* __restrict__ out) {
  const int H8 = H / 8;
  const size_t total = (size_t)B * H8;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H8), h0 = (int)(i - (size_t)b * H8) * 8;
    const T* row = e + (size_t)b * Tn * H + h0;
    float m[8];
    load8(row, m);
    for (int t = 1; t < Tn; ++t) {
      float x[8];
      load8(row + (size_t)t * H, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], x[j]);
    }
    store8(out + (size_t)b * H + h0, m);
  }
}

// one CTA per row; z f32 [B,H], v [B,H] (T), out [B,H] (T)
template <typename T>
__global__ void __launch_bounds__(256)
softmax_mul_kernel(const float* __restrict__ z, const T* __restrict__ v, int H, T* __restrict__ out) {
  __shared__ float red[8];
  __shared__ float bcast;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* zr = z + (size_t)b * H;
  float m = -INFINITY;
  for (int h = tid; h < H; h += 256) m = fmaxf(m, zr[h]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (tid == 0) { float t = red[0]; for (int w = 1; w < 8; ++w) t = fmaxf(t, red[w]); bcast = t; }
  __syncthreads();
  m = bcast;
  float s = 0.f;
  for (int h = tid; h < H; h += 256) s += expf(zr[h] - m);
  s = warp_sum(s);
  __syncthreads();
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (tid == 0) { float t = 0.f; for (int w = 0; w < 8; ++w) t += red[w]; bcast = t; }
  __syncthreads();
  const float inv = 1.f / bcast;
  for (int h = tid; h < H; h += 256)
    out[(size_t)b * H + h] = Elem<T>::from_f(expf(zr[h] - m) * inv * Elem<T>::to_f(v[(size_t)b * H + h]));
}

static int grid_for(size_t total) {
  size_t g = (total + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

// dst += src (the sum of the relation branches, encoder.py:257,264)
template <typename T>
__global__ void __launch_bounds__(256) add_inplace_kernel(T* __restrict__ dst, const T* __restrict__ src, size_t n8) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    float a[8], b[8];
    load8(dst + i * 8, a);
    load8(src + i * 8, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    store8(dst + i * 8, a);
  }
}

int add_inplace(void* dst, const void* src, size_t n, int dtype, cudaStream_t s) {
  VQA_REQUIRE(n % 8 == 0, "add_inplace: n=%zu must be a multiple of 8", n);
  if (n == 0) return VQA_OK;
  VQA_REQUIRE(dst && src, "add_inplace: NULL pointer");
  const int grid = grid_for(n / 8);
  if (dtype == VQA_BF16) add_inplace_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((__nv_bfloat16*)dst, (const __nv_bfloat16*)src, n / 8);
  else add_inplace_kernel<float><<<grid, 256, 0, s>>>((float*)dst, (const float*)src, n / 8);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

int attention_logits(const void* proj, int ldp, const float* q, int ldq, const float* w, int B, int K, int Hd, int mode,
                     int dtype, float* out, cudaStream_t s) {
  VQA_REQUIRE(B >= 0 && K >= 1 && Hd >= 8 && Hd % 8 == 0 && ldp % 8 == 0 && ldq % 8 == 0 && (mode == 0 || mode == 1),
              "attention_logits: bad dims B=%d K=%d Hd=%d ldp=%d ldq=%d mode=%d", B, K, Hd, ldp, ldq, mode);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(proj && q && w && out, "attention_logits: NULL pointer");
  const int rows = B * K;
  const int grid = grid_for((size_t)rows * 32);
  if (dtype == VQA_BF16)
    attention_logits_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)proj, ldp, q, ldq, w, rows, K, Hd,
                                                                mode, out);
  else
    attention_logits_kernel<float><<<grid, 256, 0, s>>>((const float*)proj, ldp, q, ldq, w, rows, K, Hd, mode, out);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

int caption_gate_scale(const void* out_w, const float* p, const float* r, int B, int T, int H, int dtype,
                       void* in2, float* a_out, cudaStream_t s) {
  VQA_REQUIRE(B >= 0 && T >= 1 && H >= 8 && H % 8 == 0, "caption_gate_scale: bad dims B=%d T=%d H=%d", B, T, H);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(out_w && p && r && in2, "caption_gate_scale: NULL pointer");
  const int grid = grid_for((size_t)B * (H / 8));
  if (dtype == VQA_BF16)
    caption_gate_scale_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)out_w, p, r, B, T, H,
                                                                  (__nv_bfloat16*)in2, a_out);
  else
    caption_gate_scale_kernel<float><<<grid, 256, 0, s>>>((const float*)out_w, p, r, B, T, H, (float*)in2, a_out);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

int seq_max(const void* e, int B, int T, int H, int dtype, void* out, cudaStream_t s) {
  VQA_REQUIRE(B >= 0 && T >= 1 && H >= 8 && H % 8 == 0, "seq_max: bad dims B=%d T=%d H=%d", B, T, H);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(e && out, "seq_max: NULL pointer");
  const int grid = grid_for((size_t)B * (H / 8));
  if (dtype == VQA_BF16)
    seq_max_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)e, B, T, H, (__nv_bfloat16*)out);
  else
    seq_max_kernel<float><<<grid, 256, 0, s>>>((const float*)e, B, T, H, (float*)out);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

int softmax_mul(const float* z, const void* v, int B, int H, int dtype, void* out, cudaStream_t s) {
  VQA_REQUIRE(B >= 0 && H >= 1, "softmax_mul: bad dims B=%d H=%d", B, H);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(z && v && out, "softmax_mul: NULL pointer");
  if (dtype == VQA_BF16)
    softmax_mul_kernel<__nv_bfloat16><<<B, 256, 0, s>>>(z, (const __nv_bfloat16*)v, H, (__nv_bfloat16*)out);
  else
    softmax_mul_kernel<float><<<B, 256, 0, s>>>(z, (const float*)v, H, (float*)out);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
