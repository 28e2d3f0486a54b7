// Training step of the Up-Down VQA path (BASELINE config 4): forward with saved activations,
// BCE loss, full backward down to every parameter gradient — one C call, ~120 launches.
//
// Reference semantics: Wrapper.get_loss (wrapper.py:76-105) = forward (encoder.py:146-181,
// attention.py:68-86, predictor.py:81-93) + instance_bce_with_logits (wrapper.py:25-29), followed
// by loss.backward() in train.py:108.  What torch autograd derives for that graph is restated
// here by hand:
//   * every weight-normed layer W_eff = v·g/‖v‖_F (modules.py:38; dim=None → scalar g):
//       dX = s·dY·v      dW_eff = dYᵀ·X      db = Σ_m dY
//       dv = s·(dW_eff − (Σ dW_eff⊙v / ‖v‖²)·v)      dg = Σ dW_eff⊙v / ‖v‖
//     The two GEMMs are vqa_linear's trans_w / trans_a forms (tcgen05, MN-major operands).
//   * GRU: back-propagation through time over the T steps with saved gates (torch nn.GRU
//     semantics, modules.py:153); the weight gradients are two GEMMs over all T·B rows.
//   * dropout (attention.py:74, modules.py:45; train mode only) uses a counter-based hash so the
//     mask is recomputed, not stored; p = 0 reproduces the reference's gradients exactly.
// Master parameters and gradients are f32 under the reference's names; operands are re-cast to
// the compute dtype every step (the optimiser changed them) and s = g/‖v‖ is re-evaluated on the
// device.
#include <stdlib.h>

#include "common.cuh"

namespace vqa {

int gru_pair(const void*, int, int, int, int, const void*, const void*, const float*, void*, int*, float*, void*,
             void*, const GruTrainSave*, int sm_limit, const GruTokenTable*, cudaStream_t);

int attention_pool(const float*, int, float, const void*, int, int, int, int, float*, void*, void*, cudaStream_t);
int cast_f32_to_bf16(const float*, void*, size_t, cudaStream_t);

namespace train {

enum { L_WV = 0, L_WQ, L_LIN, L_QNET, L_VNET, L_C0, L_C1, NL };
constexpr int SVEC = 4096;                 // length of the per-layer uniform scale vectors
constexpr int PARTS = 256;                 // partial sums per layer in the norm / dot reductions (fixed order: deterministic)

struct WnTable {
  const float* v[NL]; const float* g[NL]; float* dW[NL]; float* dg[NL];
  unsigned long long n[NL];
};

__device__ __forceinline__ float block_sum_256(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += red[w];
  return s;
}

// keep-mask of element idx under seed: uniform 24-bit hash >= p
__device__ __forceinline__ float keep_scale(unsigned long long seed, unsigned long long idx, float p, float inv_keep) {
  if (p <= 0.f) return 1.f;
  unsigned long long x = idx * 0x9E3779B97F4A7C15ull + seed;
  x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32; x *= 0xD6E8FEB86659FD93ull; x ^= x >> 32;
  const float u = (float)((unsigned)(x >> 40)) * (1.f / 16777216.f);
  return u >= p ? inv_keep : 0.f;
}

// ---- weight-norm scalars -------------------------------------------------------------------
// part[l][blockIdx.x] = Σ a⊙b over a strided slice (a = v, b = v for norms; a = dW, b = v for dots)
__global__ void __launch_bounds__(256) wn_dot_partials_kernel(WnTable t, int use_dw, float* __restrict__ part) {
  __shared__ float red[8];
  const int l = blockIdx.y;
  const float* a = use_dw ? t.dW[l] : t.v[l];
  const float* b = t.v[l];
  const unsigned long long n = t.n[l];
  float s = 0.f;
  if ((n & 3ull) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) {
    // streaming form: 16-byte loads, four independent chains per thread (the scalar loop ran at 0.6 TB/s)
    const unsigned long long n4 = n >> 2, stride = 256ull * gridDim.x;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n4; i += 4 * stride) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const unsigned long long j = i + u * stride;
        if (j < n4) {
          const float4 x = __ldg(a4 + j), y = __ldg(b4 + j);
          acc[u] = fmaf(x.x, y.x, fmaf(x.y, y.y, fmaf(x.z, y.z, fmaf(x.w, y.w, acc[u]))));
        }
      }
    }
    s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
  } else {
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += 256ull * gridDim.x) s = fmaf(a[i], b[i], s);
  }
  s = block_sum_256(s, red);
  if (threadIdx.x == 0) part[l * PARTS + blockIdx.x] = s;
}

// one block per layer: ‖v‖², s = g/‖v‖, uniform scale vectors, effective logit vector
__global__ void __launch_bounds__(256) wn_finalize_kernel(WnTable t, const float* __restrict__ part, float* __restrict__ scal,
                                                          float* __restrict__ svec, float* __restrict__ wlin_eff, int H,
                                                          float inv_keep_cls) {
  const int l = blockIdx.x;
  __shared__ float s_sh;
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < PARTS; ++i) acc += (double)part[l * PARTS + i];
    const float norm2 = (float)acc;
    const float s = t.g[l][0] / sqrtf(norm2);
    scal[l] = s; scal[NL + l] = norm2;
    s_sh = s;
  }
  __syncthreads();
  const float s = s_sh;
  for (int i = threadIdx.x; i < SVEC; i += 256) svec[(size_t)l * SVEC + i] = s;
  if (l == L_C1) for (int i = threadIdx.x; i < SVEC; i += 256) svec[(size_t)NL * SVEC + i] = s * inv_keep_cls;
  if (l == L_LIN) for (int i = threadIdx.x; i < H; i += 256) wlin_eff[i] = t.v[l][i] * s;
}

// one block per layer: dot = Σ dW⊙v → dg = dot/‖v‖, coef = dot/‖v‖²
__global__ void __launch_bounds__(32) wn_backward_finalize_kernel(WnTable t, const float* __restrict__ part,
                                                                  const float* __restrict__ scal, float* __restrict__ coef) {
  const int l = blockIdx.x;
  if (threadIdx.x == 0) {
    double acc = 0.0;
    for (int i = 0; i < PARTS; ++i) acc += (double)part[l * PARTS + i];
    const float norm2 = scal[NL + l];
    t.dg[l][0] = (float)(acc / sqrt((double)norm2));
    coef[l] = (float)(acc / (double)norm2);
  }
}

// dv = s·(dW − coef·v), in place over the raw dW_eff the GEMM wrote
__global__ void __launch_bounds__(256) wn_backward_apply_kernel(WnTable t, const float* __restrict__ scal, const float* __restrict__ coef) {
  const int l = blockIdx.y;
  const float s = scal[l], c = coef[l];
  float* dW = t.dW[l];
  const float* v = t.v[l];
  const unsigned long long n = t.n[l];
  if ((n & 3ull) == 0 && ((reinterpret_cast<uintptr_t>(dW) | reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    float4* d4 = reinterpret_cast<float4*>(dW);
    const float4* v4 = reinterpret_cast<const float4*>(v);
    for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < (n >> 2); i += 256ull * gridDim.x) {
      float4 d = d4[i];
      const float4 x = __ldg(v4 + i);
      d.x = s * (d.x - c * x.x); d.y = s * (d.y - c * x.y); d.z = s * (d.z - c * x.z); d.w = s * (d.w - c * x.w);
      d4[i] = d;
    }
    return;
  }
  for (unsigned long long i = blockIdx.x * 256ull + threadIdx.x; i < n; i += 256ull * gridDim.x) dW[i] = s * (dW[i] - c * v[i]);
}

// ---- operand preparation -------------------------------------------------------------------
// dst[r, 0..cols_out) = src[r, 0..cols_in) zero padded; T = compute dtype
template <typename T>
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ src, int rows, int cols_in, int cols_out,
                                                       T* __restrict__ dst) {
  const size_t total = (size_t)rows * cols_out;
  if (cols_in == cols_out && (total & 7) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < (total >> 3); i += 256ull * gridDim.x) {
      float v[8];
      load8(src + i * 8, v);
      store8(dst + i * 8, v);
    }
    return;
  }
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int r = (int)(i / cols_out), c = (int)(i - (size_t)r * cols_out);
    dst[i] = Elem<T>::from_f(c < cols_in ? src[(size_t)r * cols_in + c] : 0.f);
  }
}

// gate-interleaved packing of the GRU weights for the persistent kernel (same layout as engine.pack_gru):
// dst row j*192 + g*64 + u  <-  src row g*H + 64*j + u (zero padded to cols_out); bias [b_ir+b_hr | b_iz+b_hz | b_in | b_hn]
template <typename T>
__global__ void __launch_bounds__(256) pack_gru_kernel(const float* __restrict__ src, int H, int cols_in, int cols_out,
                                                       T* __restrict__ dst) {
  const size_t total = (size_t)3 * H * cols_out;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int p = (int)(i / cols_out), c = (int)(i - (size_t)p * cols_out);
    const int j = p / 192, g = (p - j * 192) / 64, u = p - j * 192 - g * 64;
    const int r = g * H + j * 64 + u;
    dst[i] = Elem<T>::from_f(c < cols_in ? src[(size_t)r * cols_in + c] : 0.f);
  }
}
__global__ void __launch_bounds__(256) pack_gru_bias_kernel(const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H,
                                                            float* __restrict__ out) {
  for (int i = blockIdx.x * 256 + threadIdx.x; i < 4 * H; i += 256 * gridDim.x) {
    const int g = i / H, j = i - g * H;
    out[i] = g < 2 ? b_ih[g * H + j] + b_hh[g * H + j] : (g == 2 ? b_ih[2 * H + j] : b_hh[2 * H + j]);
  }
}

// X[row, :] = emb_f32[token[row], :] (zero padded to E_pad)
template <typename T>
__global__ void __launch_bounds__(256) gather_f32_kernel(const int64_t* __restrict__ tokens, int n_rows, int E, int E_pad,
                                                         int ntoken_rows, const float* __restrict__ emb, T* __restrict__ X) {
  const size_t total = (size_t)n_rows * E_pad;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int r = (int)(i / E_pad), c = (int)(i - (size_t)r * E_pad);
    long long tok = tokens[r];
    tok = tok < 0 ? 0 : (tok >= ntoken_rows ? ntoken_rows - 1 : tok);
    X[i] = Elem<T>::from_f(c < E ? emb[(size_t)tok * E + c] : 0.f);
  }
}

// g_emb[token[row], :] += dX[row, 0..E)   (padding row ntoken_rows-1 receives no gradient)
__global__ void __launch_bounds__(256) scatter_add_kernel(const int64_t* __restrict__ tokens, int n_rows, int E, int E_pad,
                                                          int ntoken_rows, const float* __restrict__ dX, float* __restrict__ g_emb) {
  const size_t total = (size_t)n_rows * E;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int r = (int)(i / E), c = (int)(i - (size_t)r * E);
    const long long tok = tokens[r];
    if (tok < 0 || tok >= ntoken_rows - 1) continue;
    atomicAdd(g_emb + (size_t)tok * E + c, dX[(size_t)r * E_pad + c]);
  }
}

// dst[r, 0..cols) = src[r, 0..cols) with different leading dimensions (f32)
__global__ void __launch_bounds__(256) copy_cols_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst,
                                                        int ld_dst, int rows, int cols) {
  const size_t total = (size_t)rows * cols;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (size_t)r * cols);
    dst[(size_t)r * ld_dst + c] = src[(size_t)r * ld_src + c];
  }
}

// ---- GRU ------------------------------------------------------------------------------------
// forward step t with everything the backward needs saved (gates in f32, state in f32 + compute dtype)
template <typename T>
__global__ void __launch_bounds__(256) gru_gate_train_kernel(const float* __restrict__ gi, const float* __restrict__ gh, int B, int H,
                                                             int Tlen, int t, const float* __restrict__ h_prev, float* __restrict__ h_out,
                                                             T* __restrict__ h_lp, float* __restrict__ R, float* __restrict__ Z,
                                                             float* __restrict__ N, float* __restrict__ HN) {
  const int total = B * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / H, j = i - b * H;
    const float* gir = gi + ((size_t)b * Tlen + t) * 3 * H;
    const float* ghr = gh + (size_t)b * 3 * H;
    const float r = 1.f / (1.f + expf(-(gir[j] + ghr[j])));
    const float z = 1.f / (1.f + expf(-(gir[H + j] + ghr[H + j])));
    const float hn = ghr[2 * H + j];
    const float n = tanhf(gir[2 * H + j] + r * hn);
    const float h = (1.f - z) * n + z * h_prev[i];
    R[i] = r; Z[i] = z; N[i] = n; HN[i] = hn;
    h_out[i] = h;
    h_lp[i] = Elem<T>::from_f(h);
  }
}

// backward of step t: dGI row (b*T+t) = [dr, dz, dn] pre-activation grads, dGH row b = [dr, dz, dn·r],
// dh_part = dh ⊙ z (the direct path to h_{t-1}; the GEMM adds dGH·W_hh on top)
template <typename T>
__global__ void __launch_bounds__(256) gru_gate_backward_kernel(const float* __restrict__ dh, const float* __restrict__ R,
                                                                const float* __restrict__ Z, const float* __restrict__ N,
                                                                const float* __restrict__ HN, const float* __restrict__ h_prev,
                                                                int B, int H, int Tlen, int t, T* __restrict__ dGI,
                                                                T* __restrict__ dGH, float* __restrict__ dh_part) {
  const int total = B * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / H, j = i - b * H;
    const float d = dh[i], r = R[i], z = Z[i], n = N[i], hn = HN[i];
    const float dn_pre = d * (1.f - z) * (1.f - n * n);
    const float dz_pre = d * (h_prev[i] - n) * z * (1.f - z);
    const float dr_pre = dn_pre * hn * r * (1.f - r);
    T* gi = dGI + ((size_t)b * Tlen + t) * 3 * H;
    T* gh = dGH + (size_t)b * 3 * H;
    gi[j] = Elem<T>::from_f(dr_pre); gi[H + j] = Elem<T>::from_f(dz_pre); gi[2 * H + j] = Elem<T>::from_f(dn_pre);
    gh[j] = Elem<T>::from_f(dr_pre); gh[H + j] = Elem<T>::from_f(dz_pre); gh[2 * H + j] = Elem<T>::from_f(dn_pre * r);
    dh_part[i] = d * z;
  }
}

// ---- attention ------------------------------------------------------------------------------
// logit[b,k] = Σ_h Vp[b,k,h]·Qp[b,h]·wl[h]·keep   (attention.py:72-75 with the projection stored)
template <typename T>
__global__ void __launch_bounds__(256) att_logit_kernel(const T* __restrict__ Vp, const float* __restrict__ qq, int ld_qq,
                                                        const float* __restrict__ wl, int K, int H, float p, float inv_keep,
                                                        unsigned long long seed, float* __restrict__ logit) {
  extern __shared__ float qw[];                      // Qp[b,:] ⊙ wl
  const int b = blockIdx.x;
  for (int h = threadIdx.x; h < H; h += 256) qw[h] = qq[(size_t)b * ld_qq + h] * wl[h];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int k = warp; k < K; k += 8) {
    const size_t row = (size_t)b * K + k;
    const T* vr = Vp + row * H;
    float s = 0.f;
    for (int h = lane; h < H; h += 32) s = fmaf(Elem<T>::to_f(vr[h]) * keep_scale(seed, row * H + h, p, inv_keep), qw[h], s);
    s = warp_sum(s);
    if (lane == 0) logit[row] = s;
  }
}

// datt[b,k] = dvsum[b]·x[b,k]; dlogit = att ⊙ (datt − Σ_j att_j datt_j)   (encoder.py:166 + softmax backward)
template <typename T>
__global__ void __launch_bounds__(256) pool_backward_kernel(const float* __restrict__ dvsum, const T* __restrict__ x,
                                                            const float* __restrict__ att, int K, int V, float* __restrict__ dlogit) {
  __shared__ float datt[64];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float* dv = dvsum + (size_t)b * V;
  for (int k = warp; k < K; k += 8) {
    const T* xr = x + ((size_t)b * K + k) * V;
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s = fmaf(dv[c], Elem<T>::to_f(xr[c]), s);
    s = warp_sum(s);
    if (lane == 0) datt[k] = s;
  }
  __syncthreads();
  if (threadIdx.x < K) {
    float dot = 0.f;
    for (int j = 0; j < K; ++j) dot = fmaf(att[(size_t)b * K + j], datt[j], dot);
    const int k = threadIdx.x;
    dlogit[(size_t)b * K + k] = att[(size_t)b * K + k] * (datt[k] - dot);
  }
}

// in place Vp → dVp (pre-activation grad of W_v's layer); dQp → dqq[:, :H] (pre-activation, ⊙1[Qp>0]);
// per-image partials of the logit layer's weight / bias gradient
template <typename T>
__global__ void __launch_bounds__(256) att_backward_kernel(T* __restrict__ Vp, const float* __restrict__ qq, int ld_qq,
                                                           const float* __restrict__ wl, const float* __restrict__ dlogit, int K, int H,
                                                           float p, float inv_keep, unsigned long long seed, T* __restrict__ dqq,
                                                           int ld_dqq, float* __restrict__ dwl_part, float* __restrict__ dbl_part) {
  __shared__ float dl[64];
  const int b = blockIdx.x;
  if (threadIdx.x < K) dl[threadIdx.x] = dlogit[(size_t)b * K + threadIdx.x];
  __syncthreads();
  for (int h = threadIdx.x; h < H; h += 256) {
    const float q = qq[(size_t)b * ld_qq + h], w = wl[h];
    float dq = 0.f, dw = 0.f;
    for (int k = 0; k < K; ++k) {
      const size_t idx = ((size_t)b * K + k) * H + h;
      const float vp = Elem<T>::to_f(Vp[idx]);
      const float m = keep_scale(seed, idx, p, inv_keep);
      const float g = dl[k] * m;
      dq = fmaf(g * w, vp, dq);
      dw = fmaf(g * q, vp, dw);
      Vp[idx] = Elem<T>::from_f(vp > 0.f ? g * w * q : 0.f);
    }
    dqq[(size_t)b * ld_dqq + h] = Elem<T>::from_f(q > 0.f ? dq : 0.f);
    dwl_part[(size_t)b * H + h] = dw;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int k = 0; k < K; ++k) s += dl[k];
    dbl_part[b] = s;
  }
}

// ---- predictor elementwise -------------------------------------------------------------------
// joint = vn ⊙ qn   (predictor.py:91)
template <typename T>
__global__ void __launch_bounds__(256) joint_kernel(const T* __restrict__ vn, const float* __restrict__ qn, int ld_qn, int B, int H,
                                                    T* __restrict__ joint) {
  const int total = B * H;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += 256 * gridDim.x) {
    const int b = i / H, j = i - b * H;
    joint[i] = Elem<T>::from_f(Elem<T>::to_f(vn[i]) * qn[(size_t)b * ld_qn + j]);
  }
}
// dvn_pre = djoint ⊙ qn ⊙ 1[vn>0];  dqn_pre = djoint ⊙ vn ⊙ 1[qn>0] → dqq[:, H:]
template <typename T>
__global__ void __launch_bounds__(256) joint_backward_kernel(const float* __restrict__ dj, const T* __restrict__ vn,
                                                             const float* __restrict__ qn, int ld_qn, int B, int H,
                                                             T* __restrict__ dvn, T* __restrict__ dqn, int ld_dqn) {
  const int total = B * H;
  for (int i = blockIdx.x * 256 + threadIdx.x; i < total; i += 256 * gridDim.x) {
    const int b = i / H, j = i - b * H;
    const float v = Elem<T>::to_f(vn[i]), q = qn[(size_t)b * ld_qn + j], d = dj[i];
    dvn[i] = Elem<T>::from_f(v > 0.f ? d * q : 0.f);
    dqn[(size_t)b * ld_dqn + j] = Elem<T>::from_f(q > 0.f ? d * v : 0.f);
  }
}
// hid ← hid ⊙ keep/(1-p)   (modules.py:45, nn.Dropout(inplace=True) after the ReLU)
template <typename T>
__global__ void __launch_bounds__(256) dropout_kernel(T* __restrict__ x, size_t n, float p, float inv_keep, unsigned long long seed) {
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += 256ull * gridDim.x)
    x[i] = Elem<T>::from_f(Elem<T>::to_f(x[i]) * keep_scale(seed, i, p, inv_keep));
}

// ---- loss ------------------------------------------------------------------------------------
// loss = mean_{b,a} BCEWithLogits(z, t) · A (wrapper.py:25-29) = Σ / B;  dz = (σ(z) − t)/B ⊙ 1[z>0]
// (z is the classifier's ReLU output, modules.py:55: its pre-activation gradient is masked here)
template <typename T>
__global__ void __launch_bounds__(256) bce_loss_grad_kernel(const float* __restrict__ z, const float* __restrict__ target, int B, int A,
                                                            int ldd, T* __restrict__ dz, float* __restrict__ part) {
  __shared__ float red[8];
  const size_t total = (size_t)B * ldd;
  const float invB = 1.f / (float)B;
  float acc = 0.f;
  for (size_t i = blockIdx.x * 256ull + threadIdx.x; i < total; i += 256ull * gridDim.x) {
    const int b = (int)(i / ldd), a = (int)(i - (size_t)b * ldd);
    float d = 0.f;
    if (a < A) {
      const float x = z[(size_t)b * A + a], t = target[(size_t)b * A + a];
      acc += fmaxf(x, 0.f) - x * t + log1pf(expf(-fabsf(x)));
      const float sg = 1.f / (1.f + expf(-x));
      d = x > 0.f ? (sg - t) * invB : 0.f;
    }
    dz[i] = Elem<T>::from_f(d);
  }
  acc = block_sum_256(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = acc;
}
__global__ void __launch_bounds__(32) loss_finalize_kernel(const float* __restrict__ part, int n, int B, float* __restrict__ loss) {
  double acc = 0.0;                                   // lane-strided partial sums, then a fixed-order tree: deterministic
  for (int i = threadIdx.x; i < n; i += 32) acc += (double)part[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) loss[0] = (float)(acc / (double)B);
}

// ---- column sums (bias gradients): out[n] = Σ_m x[m,n], deterministic two-pass in one launch ----
constexpr int CS_CHUNKS = 32;
template <typename T>
__global__ void __launch_bounds__(256) colsum_kernel(const T* __restrict__ x, int M, int N, int ld, float* __restrict__ part,
                                                     unsigned int* __restrict__ tickets, float* __restrict__ out) {
  __shared__ float sm[4][64];
  __shared__ bool last;
  const int c = threadIdx.x & 63, rl = threadIdx.x >> 6;
  const int n = blockIdx.x * 64 + c;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
  float s = 0.f;
  if (n < N) for (int r = r0 + rl; r < r1; r += 4) s += Elem<T>::to_f(x[(size_t)r * ld + n]);
  sm[rl][c] = s;
  __syncthreads();
  if (rl == 0 && n < N) part[(size_t)blockIdx.y * N + n] = sm[0][c] + sm[1][c] + sm[2][c] + sm[3][c];
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = (atomicAdd(tickets + blockIdx.x, 1u) == gridDim.y - 1);
  __syncthreads();
  if (last) {
    __threadfence();
    if (rl == 0 && n < N) {
      float t = 0.f;
      for (int y = 0; y < (int)gridDim.y; ++y) t += __ldcg(part + (size_t)y * N + n);
      out[n] = t;
    }
    if (threadIdx.x == 0) tickets[blockIdx.x] = 0;          // ready for the next launch
  }
}

}  // namespace train

// =================================================================================================
using namespace train;

static int linear_dispatch_t(const vqa_linear_args& a, cudaStream_t s) {
  if (a.dtype == VQA_BF16) {
    static int force = -1;
    if (force < 0) { const char* e = getenv("VQA_B200_FORCE_SIMT"); force = (e && e[0] == '1') ? 1 : 0; }
    if (!force) return linear_tc(a, s);
  }
  return linear_simt(a, s);
}

struct TrainWs {
  // compute-dtype operands
  void *Wv, *Wqq, *Wvn, *Wc0, *Wc1, *w_ih, *w_hh;
  float *scal, *svec, *wlin, *bqq, *part, *coef;
  void* X; float* GI; float* GH; float* Hs; void* Hlp; float *R, *Z, *N, *HN;
  void *wx_packed, *wh_packed; float* bias_packed; int* gru_counter;       // persistent forward GRU (bf16)
  float* qq; void* Vp; float* logit; float* att; void* vsum; void* vn; void* joint; void* hid;
  void* dlogits; void* dhid; float* djoint; void* dvn; void* dqq; float* dvsum; float* dlogit;
  float* dwl_part; float* dbl_part; float* dh0; float* dh1; float* dh_part; void* dGI; void* dGH; float* dX; float* dwih_pad;
  float* cs_part; unsigned int* tickets; float* loss_part; float* dwl; float* dbl;
  size_t bytes;
};

static int round8(int x) { return (x + 7) / 8 * 8; }
static int e_pad(int E) { return (E + 63) / 64 * 64; }

static TrainWs carve_train(const vqa_train_args& a, void* base) {
  TrainWs w{};
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t n) { void* r = p ? p + off : nullptr; off += align_up(n, 256); return r; };
  const size_t es = elem_size(a.dtype);
  const size_t B = a.B, K = a.K, V = a.V, H = a.H, A = a.A, T = a.T, Ep = e_pad(a.E), ldA = round8(a.A);
  w.Wv = take(H * V * es); w.Wqq = take(2 * H * H * es); w.Wvn = take(H * V * es); w.Wc0 = take(2 * H * H * es);
  w.Wc1 = take(A * 2 * H * es); w.w_ih = take(3 * H * Ep * es); w.w_hh = take(3 * H * H * es);
  w.scal = (float*)take(2 * NL * 4); w.svec = (float*)take((size_t)(NL + 1) * SVEC * 4); w.wlin = (float*)take(H * 4);
  w.bqq = (float*)take(2 * H * 4); w.part = (float*)take((size_t)NL * PARTS * 4); w.coef = (float*)take(NL * 4);
  w.X = take(B * T * Ep * es); w.GI = (float*)take(B * T * 3 * H * 4); w.GH = (float*)take(B * 3 * H * 4);
  w.Hs = (float*)take((T + 1) * B * H * 4); w.Hlp = take((T + 1) * B * H * es);
  w.R = (float*)take(T * B * H * 4); w.Z = (float*)take(T * B * H * 4); w.N = (float*)take(T * B * H * 4); w.HN = (float*)take(T * B * H * 4);
  w.wx_packed = take(3 * H * Ep * es); w.wh_packed = take(3 * H * H * es); w.bias_packed = (float*)take(4 * H * 4);
  w.gru_counter = (int*)take(256);
  w.qq = (float*)take(B * 2 * H * 4); w.Vp = take(B * K * H * es); w.logit = (float*)take(B * K * 4); w.att = (float*)take(B * K * 4);
  w.vsum = take(B * V * es); w.vn = take(B * H * es); w.joint = take(B * H * es); w.hid = take(B * 2 * H * es);
  w.dlogits = take(B * ldA * es); w.dhid = take(B * 2 * H * es); w.djoint = (float*)take(B * H * 4); w.dvn = take(B * H * es);
  w.dqq = take(B * 2 * H * es); w.dvsum = (float*)take(B * V * 4); w.dlogit = (float*)take(B * K * 4);
  w.dwl_part = (float*)take(B * H * 4); w.dbl_part = (float*)take(B * 4);
  w.dh0 = (float*)take(B * H * 4); w.dh1 = (float*)take(B * H * 4); w.dh_part = (float*)take(B * H * 4);
  w.dGI = take(B * T * 3 * H * es); w.dGH = take(T * B * 3 * H * es); w.dX = (float*)take(B * T * Ep * 4);
  w.dwih_pad = (float*)take(3 * H * Ep * 4);
  const size_t maxN = ldA > 3 * H ? ldA : 3 * H;
  w.cs_part = (float*)take((size_t)CS_CHUNKS * maxN * 4); w.tickets = (unsigned int*)take(((maxN + 63) / 64) * 4);
  w.loss_part = (float*)take(1024 * 4); w.dwl = (float*)take(H * 4); w.dbl = (float*)take(256);
  w.bytes = off;
  return w;
}

static int grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  const size_t cap = (size_t)sm_count() * 8;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

template <typename T>
static int train_step_t(const vqa_train_args& a, const TrainWs& w, cudaStream_t s) {
  const int B = a.B, K = a.K, V = a.V, H = a.H, A = a.A, Tn = a.T, E = a.E, Ep = e_pad(a.E), ldA = round8(a.A);
  const bool bf16 = a.dtype == VQA_BF16;
  int rc;
  // the ticket counters of colsum must start at zero (they reset themselves afterwards)
  const size_t maxN = (size_t)(ldA > 3 * H ? ldA : 3 * H);
  VQA_CUDA_CHECK(cudaMemsetAsync(w.tickets, 0, ((maxN + 63) / 64) * 4, s));

  // ---------------- operand preparation ----------------
  WnTable tb{};
  const int out_dim[NL] = {H, H, 1, H, H, 2 * H, A};
  const int in_dim[NL] = {V, H, H, H, V, H, 2 * H};
  for (int l = 0; l < NL; ++l) {
    tb.v[l] = a.p_v[l]; tb.g[l] = a.p_g[l]; tb.dW[l] = a.g_v[l]; tb.dg[l] = a.g_g[l];
    tb.n[l] = (unsigned long long)out_dim[l] * in_dim[l];
  }
  const float inv_keep_att = 1.f / (1.f - a.dropout_att), inv_keep_cls = 1.f / (1.f - a.dropout_cls);
  wn_dot_partials_kernel<<<dim3(PARTS, NL), 256, 0, s>>>(tb, 0, w.part);
  VQA_LAUNCH_CHECK();
  wn_finalize_kernel<<<NL, 256, 0, s>>>(tb, w.part, w.scal, w.svec, w.wlin, H, inv_keep_cls);
  VQA_LAUNCH_CHECK();
  auto svec = [&](int l) { return w.svec + (size_t)l * SVEC; };
  auto cast = [&](const float* src, void* dst, size_t rows, int cols) -> int {
    cast_pad_kernel<T><<<grid_for(rows * cols), 256, 0, s>>>(src, (int)rows, cols, cols, (T*)dst);
    VQA_LAUNCH_CHECK();
    return VQA_OK;
  };
  if ((rc = cast(a.p_v[L_WV], w.Wv, H, V))) return rc;
  if ((rc = cast(a.p_v[L_WQ], w.Wqq, H, H))) return rc;
  if ((rc = cast(a.p_v[L_QNET], (char*)w.Wqq + (size_t)H * H * sizeof(T), H, H))) return rc;
  if ((rc = cast(a.p_v[L_VNET], w.Wvn, H, V))) return rc;
  if ((rc = cast(a.p_v[L_C0], w.Wc0, 2 * H, H))) return rc;
  if ((rc = cast(a.p_v[L_C1], w.Wc1, A, 2 * H))) return rc;
  if ((rc = cast(a.p_w_hh, w.w_hh, 3 * H, H))) return rc;
  cast_pad_kernel<T><<<grid_for((size_t)3 * H * Ep), 256, 0, s>>>(a.p_w_ih, 3 * H, E, Ep, (T*)w.w_ih);
  VQA_LAUNCH_CHECK();
  VQA_CUDA_CHECK(cudaMemcpyAsync(w.bqq, a.p_b[L_WQ], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  VQA_CUDA_CHECK(cudaMemcpyAsync(w.bqq + H, a.p_b[L_QNET], (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  // [s_q | s_n] for the concatenated [W_q ; q_net] GEMM lives in the spare scale-vector slot NL... use two halves
  float* sqq = w.svec + (size_t)NL * SVEC + 2 * H;          // beyond the 2H entries the c1 backward reads
  VQA_CUDA_CHECK(cudaMemcpyAsync(sqq, svec(L_WQ), (size_t)H * 4, cudaMemcpyDeviceToDevice, s));
  VQA_CUDA_CHECK(cudaMemcpyAsync(sqq + H, svec(L_QNET), (size_t)H * 4, cudaMemcpyDeviceToDevice, s));

  auto lin = [&](vqa_linear_args l) -> int { l.dtype = a.dtype; if (l.mul_row_div < 1) l.mul_row_div = 1; if (l.add_row_div < 1) l.add_row_div = 1; return linear_dispatch_t(l, s); };
  const int dt = a.dtype;

  // ---------------- forward ----------------
  gather_f32_kernel<T><<<grid_for((size_t)B * Tn * Ep), 256, 0, s>>>(a.d_tokens, B * Tn, E, Ep, a.ntoken_rows, a.p_emb, (T*)w.X);
  VQA_LAUNCH_CHECK();
  VQA_CUDA_CHECK(cudaMemsetAsync(w.Hs, 0, (size_t)B * H * 4, s));
  VQA_CUDA_CHECK(cudaMemsetAsync(w.Hlp, 0, (size_t)B * H * sizeof(T), s));
  const size_t BH = (size_t)B * H;
  // Forward GRU.  bf16: ONE launch of the persistent pair kernel (gru_pair.cu) in its training form — states written
  // time-major into Hs / Hlp, gates r, z, n and W_hn·h + b_hn saved per step; the input projection is fused, so the
  // [B*T,3H] gi GEMM and the 2T per-step launches disappear.  fp32 (and devices that cannot hold the pairs): per step.
  int gru_rc = VQA_ERR_UNSUPPORTED;
  if (bf16 && H % 64 == 0 && !getenv("VQA_B200_TRAIN_GRU_STEPWISE") && !getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR")) {   // (ncu: see api.cu)
    pack_gru_kernel<T><<<grid_for((size_t)3 * H * Ep), 256, 0, s>>>(a.p_w_ih, H, E, Ep, (T*)w.wx_packed);
    VQA_LAUNCH_CHECK();
    pack_gru_kernel<T><<<grid_for((size_t)3 * H * H), 256, 0, s>>>(a.p_w_hh, H, H, H, (T*)w.wh_packed);
    VQA_LAUNCH_CHECK();
    pack_gru_bias_kernel<<<(4 * H + 255) / 256, 256, 0, s>>>(a.p_b_ih, a.p_b_hh, H, w.bias_packed);
    VQA_LAUNCH_CHECK();
    const GruTrainSave save{w.R, w.Z, w.N, w.HN, w.Hs + BH};
    gru_rc = gru_pair(w.X, B, Tn, H, Ep, w.wx_packed, w.wh_packed, w.bias_packed, nullptr, w.gru_counter, nullptr, nullptr,
                      (T*)w.Hlp + BH, &save, 0, nullptr, s);
    if (gru_rc != VQA_OK && gru_rc != VQA_ERR_UNSUPPORTED) return gru_rc;
  }
  if (gru_rc == VQA_ERR_UNSUPPORTED) {
    {
      vqa_linear_args l{};
      l.d_A = w.X; l.lda = Ep; l.d_W = w.w_ih; l.ldw = Ep; l.M = B * Tn; l.N = 3 * H; l.K = Ep; l.d_bias = a.p_b_ih;
      l.d_out = w.GI; l.ldo = 3 * H; l.out_dtype = VQA_F32;
      if ((rc = lin(l))) return rc;
    }
    for (int t = 0; t < Tn; ++t) {
      vqa_linear_args l{};
      l.d_A = (T*)w.Hlp + t * BH; l.lda = H; l.d_W = w.w_hh; l.ldw = H; l.M = B; l.N = 3 * H; l.K = H; l.d_bias = a.p_b_hh;
      l.d_out = w.GH; l.ldo = 3 * H; l.out_dtype = VQA_F32;
      if ((rc = lin(l))) return rc;
      gru_gate_train_kernel<T><<<grid_for(BH), 256, 0, s>>>(w.GI, w.GH, B, H, Tn, t, w.Hs + t * BH, w.Hs + (t + 1) * BH,
                                                           (T*)w.Hlp + (t + 1) * BH, w.R + t * BH, w.Z + t * BH, w.N + t * BH, w.HN + t * BH);
      VQA_LAUNCH_CHECK();
    }
  }
  const T* hT = (T*)w.Hlp + (size_t)Tn * BH;
  {  // qq = ReLU([W_q;q_net] h) f32 [B,2H]
    vqa_linear_args l{};
    l.d_A = hT; l.lda = H; l.d_W = w.Wqq; l.ldw = H; l.M = B; l.N = 2 * H; l.K = H; l.d_scale = sqq; l.d_bias = w.bqq; l.relu = 1;
    l.d_out = w.qq; l.ldo = 2 * H; l.out_dtype = VQA_F32;
    if ((rc = lin(l))) return rc;
  }
  {  // Vp = ReLU(W_v x) stored (the backward needs its sign and values)
    vqa_linear_args l{};
    l.d_A = a.d_img; l.lda = V; l.d_W = w.Wv; l.ldw = V; l.M = B * K; l.N = H; l.K = V; l.d_scale = svec(L_WV); l.d_bias = a.p_b[L_WV];
    l.relu = 1; l.d_out = w.Vp; l.ldo = H; l.out_dtype = dt;
    if ((rc = lin(l))) return rc;
  }
  att_logit_kernel<T><<<B, 256, H * 4, s>>>((const T*)w.Vp, w.qq, 2 * H, w.wlin, K, H, a.dropout_att, inv_keep_att, a.seed, w.logit);
  VQA_LAUNCH_CHECK();
  float b_lin_host = 0.f;   // the logit bias shifts every region equally: the softmax ignores it, its gradient is Σ dlogit = 0
  if ((rc = attention_pool(w.logit, 1, b_lin_host, a.d_img, B, K, V, dt, w.att, w.vsum, nullptr, s))) return rc;
  {  // vn = ReLU(v_net vsum)
    vqa_linear_args l{};
    l.d_A = w.vsum; l.lda = V; l.d_W = w.Wvn; l.ldw = V; l.M = B; l.N = H; l.K = V; l.d_scale = svec(L_VNET); l.d_bias = a.p_b[L_VNET];
    l.relu = 1; l.d_out = w.vn; l.ldo = H; l.out_dtype = dt;
    if ((rc = lin(l))) return rc;
  }
  joint_kernel<T><<<grid_for(BH), 256, 0, s>>>((const T*)w.vn, w.qq + H, 2 * H, B, H, (T*)w.joint);
  VQA_LAUNCH_CHECK();
  {  // hid = dropout(ReLU(c0 joint))
    vqa_linear_args l{};
    l.d_A = w.joint; l.lda = H; l.d_W = w.Wc0; l.ldw = H; l.M = B; l.N = 2 * H; l.K = H; l.d_scale = svec(L_C0); l.d_bias = a.p_b[L_C0];
    l.relu = 1; l.d_out = w.hid; l.ldo = 2 * H; l.out_dtype = dt;
    if ((rc = lin(l))) return rc;
  }
  if (a.dropout_cls > 0.f) {
    dropout_kernel<T><<<grid_for(2 * BH), 256, 0, s>>>((T*)w.hid, 2 * BH, a.dropout_cls, inv_keep_cls, a.seed ^ 0xC15ull);
    VQA_LAUNCH_CHECK();
  }
  {  // logits = ReLU(c1 hid) f32
    vqa_linear_args l{};
    l.d_A = w.hid; l.lda = 2 * H; l.d_W = w.Wc1; l.ldw = 2 * H; l.M = B; l.N = A; l.K = 2 * H; l.d_scale = svec(L_C1); l.d_bias = a.p_b[L_C1];
    l.relu = 1; l.d_out = a.d_logits; l.ldo = A; l.out_dtype = VQA_F32;
    if ((rc = lin(l))) return rc;
  }
  const int loss_grid = grid_for((size_t)B * ldA) > 1024 ? 1024 : grid_for((size_t)B * ldA);
  bce_loss_grad_kernel<T><<<loss_grid, 256, 0, s>>>(a.d_logits, a.d_target, B, A, ldA, (T*)w.dlogits, w.loss_part);
  VQA_LAUNCH_CHECK();
  loss_finalize_kernel<<<1, 32, 0, s>>>(w.loss_part, loss_grid, B, a.d_loss);
  VQA_LAUNCH_CHECK();

  // ---------------- backward ----------------
  auto colsum = [&](const void* x, int M, int N, int ld, bool is_f32, float* out) -> int {
    dim3 g((N + 63) / 64, M < CS_CHUNKS * 4 ? 1 : CS_CHUNKS);
    if (is_f32) colsum_kernel<float><<<g, 256, 0, s>>>((const float*)x, M, N, ld, w.cs_part, w.tickets, out);
    else colsum_kernel<T><<<g, 256, 0, s>>>((const T*)x, M, N, ld, w.cs_part, w.tickets, out);
    VQA_LAUNCH_CHECK();
    return VQA_OK;
  };
  // dW_eff = dYᵀ·X  (f32 straight into the gradient buffer; weight-norm chain rule applied at the end)
  auto gemm_dW = [&](const void* dY, int ld_dy, const void* X, int ldx, int rows, int N, int Kin, float* out, int ldo) -> int {
    vqa_linear_args l{};
    l.d_A = dY; l.lda = ld_dy; l.d_W = X; l.ldw = ldx; l.M = N; l.N = Kin; l.K = rows; l.trans_a = 1; l.trans_w = 1;
    l.d_out = out; l.ldo = ldo; l.out_dtype = VQA_F32;
    return lin(l);
  };
  // dX = (dY·W)·s [+ add] [masked]
  auto gemm_dX = [&](const void* dY, int ld_dy, const void* W, int ldw, int M, int N, int Kin, const float* scale, const float* add,
                     const void* mask, int ld_mask, int mask_dtype, void* out, int ldo, int out_dtype) -> int {
    vqa_linear_args l{};
    l.d_A = dY; l.lda = ld_dy; l.d_W = W; l.ldw = ldw; l.M = M; l.N = Kin; l.K = N; l.trans_w = 1;
    l.d_scale = scale; l.d_add = add; l.ld_add = Kin; l.add_row_div = 1; l.d_mask = mask; l.ld_mask = ld_mask; l.mask_dtype = mask_dtype;
    l.d_out = out; l.ldo = ldo; l.out_dtype = out_dtype;
    return lin(l);
  };
  // classifier layer 2 (predictor.classifier.main.3)
  if ((rc = colsum(w.dlogits, B, A, ldA, false, a.g_b[L_C1]))) return rc;
  if ((rc = gemm_dW(w.dlogits, ldA, w.hid, 2 * H, B, A, 2 * H, a.g_v[L_C1], 2 * H))) return rc;
  if ((rc = gemm_dX(w.dlogits, ldA, w.Wc1, 2 * H, B, A, 2 * H, w.svec + (size_t)NL * SVEC, nullptr, w.hid, 2 * H, dt, w.dhid, 2 * H, dt))) return rc;
  // classifier layer 1 (predictor.classifier.main.0)
  if ((rc = colsum(w.dhid, B, 2 * H, 2 * H, false, a.g_b[L_C0]))) return rc;
  if ((rc = gemm_dW(w.dhid, 2 * H, w.joint, H, B, 2 * H, H, a.g_v[L_C0], H))) return rc;
  if ((rc = gemm_dX(w.dhid, 2 * H, w.Wc0, H, B, 2 * H, H, svec(L_C0), nullptr, nullptr, 0, 0, w.djoint, H, VQA_F32))) return rc;
  // joint = q ⊙ v
  T* dqq = (T*)w.dqq;
  joint_backward_kernel<T><<<grid_for(BH), 256, 0, s>>>(w.djoint, (const T*)w.vn, w.qq + H, 2 * H, B, H, (T*)w.dvn, dqq + H, 2 * H);
  VQA_LAUNCH_CHECK();
  // v_net
  if ((rc = colsum(w.dvn, B, H, H, false, a.g_b[L_VNET]))) return rc;
  if ((rc = gemm_dW(w.dvn, H, w.vsum, V, B, H, V, a.g_v[L_VNET], V))) return rc;
  if ((rc = gemm_dX(w.dvn, H, w.Wvn, V, B, H, V, svec(L_VNET), nullptr, nullptr, 0, 0, w.dvsum, V, VQA_F32))) return rc;
  // attention-weighted sum + softmax
  pool_backward_kernel<T><<<B, 256, 0, s>>>(w.dvsum, (const T*)a.d_img, w.att, K, V, w.dlogit);
  VQA_LAUNCH_CHECK();
  // logit layer, ⊙, ReLU of both projections
  att_backward_kernel<T><<<B, 256, 0, s>>>((T*)w.Vp, w.qq, 2 * H, w.wlin, w.dlogit, K, H, a.dropout_att, inv_keep_att, a.seed, dqq, 2 * H,
                                           w.dwl_part, w.dbl_part);
  VQA_LAUNCH_CHECK();
  // dwl_part is d/d(w_eff) of the logit layer: with W_eff = s·v its raw gradient is the [1,H] "dW_eff"
  if ((rc = colsum(w.dwl_part, B, H, H, true, a.g_v[L_LIN]))) return rc;
  if ((rc = colsum(w.dbl_part, B, 1, 1, true, a.g_b[L_LIN]))) return rc;
  // W_v (no gradient into the image features)
  if ((rc = colsum(w.Vp, B * K, H, H, false, a.g_b[L_WV]))) return rc;
  if ((rc = gemm_dW(w.Vp, H, a.d_img, V, B * K, H, V, a.g_v[L_WV], V))) return rc;
  // [W_q ; q_net]
  if ((rc = colsum(w.dqq, B, H, 2 * H, false, a.g_b[L_WQ]))) return rc;
  if ((rc = colsum(dqq + H, B, H, 2 * H, false, a.g_b[L_QNET]))) return rc;
  if ((rc = gemm_dW(dqq, 2 * H, hT, H, B, H, H, a.g_v[L_WQ], H))) return rc;
  if ((rc = gemm_dW(dqq + H, 2 * H, hT, H, B, H, H, a.g_v[L_QNET], H))) return rc;
  if ((rc = gemm_dX(dqq, 2 * H, w.Wqq, H, B, H, H, svec(L_WQ), nullptr, nullptr, 0, 0, w.dh1, H, VQA_F32))) return rc;
  if ((rc = gemm_dX(dqq + H, 2 * H, (T*)w.Wqq + (size_t)H * H, H, B, H, H, svec(L_QNET), w.dh1, nullptr, 0, 0, w.dh0, H, VQA_F32))) return rc;
  // weight-norm chain rule for the 7 layers, in place over the raw dW_eff.  All seven weight-normed layers are "head" layers:
  // their gradients are final HERE, before the back-propagation through time — the caller's event marks that point, so a
  // data-parallel job can all-reduce this bucket (59 of the 75.5 MB) on a side stream under the BPTT (training.py).
  wn_dot_partials_kernel<<<dim3(PARTS, NL), 256, 0, s>>>(tb, 1, w.part);
  VQA_LAUNCH_CHECK();
  wn_backward_finalize_kernel<<<NL, 32, 0, s>>>(tb, w.part, w.scal, w.coef);
  VQA_LAUNCH_CHECK();
  wn_backward_apply_kernel<<<dim3(PARTS * 4, NL), 256, 0, s>>>(tb, w.scal, w.coef);
  VQA_LAUNCH_CHECK();
  if (a.ev_head_done) VQA_CUDA_CHECK(cudaEventRecord((cudaEvent_t)a.ev_head_done, s));
  // GRU, back-propagation through time
  float* dh_cur = w.dh0;
  float* dh_nxt = w.dh1;
  T* dGH = (T*)w.dGH;
  for (int t = Tn - 1; t >= 0; --t) {
    gru_gate_backward_kernel<T><<<grid_for(BH), 256, 0, s>>>(dh_cur, w.R + t * BH, w.Z + t * BH, w.N + t * BH, w.HN + t * BH, w.Hs + t * BH, B, H,
                                                            Tn, t, (T*)w.dGI, dGH + (size_t)t * B * 3 * H, w.dh_part);
    VQA_LAUNCH_CHECK();
    if (t > 0) {
      if ((rc = gemm_dX(dGH + (size_t)t * B * 3 * H, 3 * H, w.w_hh, H, B, 3 * H, H, nullptr, w.dh_part, nullptr, 0, 0, dh_nxt, H, VQA_F32))) return rc;
      float* tmp = dh_cur; dh_cur = dh_nxt; dh_nxt = tmp;
    }
  }
  if ((rc = colsum(w.dGI, B * Tn, 3 * H, 3 * H, false, a.g_b_ih))) return rc;
  if ((rc = colsum(w.dGH, B * Tn, 3 * H, 3 * H, false, a.g_b_hh))) return rc;
  if ((rc = gemm_dW(w.dGH, 3 * H, w.Hlp, H, Tn * B, 3 * H, H, a.g_w_hh, H))) return rc;       // rows (t,b) ↔ h_{t-1}
  if ((rc = gemm_dW(w.dGI, 3 * H, w.X, Ep, B * Tn, 3 * H, Ep, w.dwih_pad, Ep))) return rc;
  copy_cols_kernel<<<grid_for((size_t)3 * H * E), 256, 0, s>>>(w.dwih_pad, Ep, a.g_w_ih, E, 3 * H, E);
  VQA_LAUNCH_CHECK();
  if ((rc = gemm_dX(w.dGI, 3 * H, w.w_ih, Ep, B * Tn, 3 * H, Ep, nullptr, nullptr, nullptr, 0, 0, w.dX, Ep, VQA_F32))) return rc;
  VQA_CUDA_CHECK(cudaMemsetAsync(a.g_emb, 0, (size_t)a.ntoken_rows * E * 4, s));
  scatter_add_kernel<<<grid_for((size_t)B * Tn * E), 256, 0, s>>>(a.d_tokens, B * Tn, E, Ep, a.ntoken_rows, w.dX, a.g_emb);
  VQA_LAUNCH_CHECK();
  (void)bf16;
  return VQA_OK;
}

size_t train_workspace_bytes(const vqa_train_args& a) { return carve_train(a, nullptr).bytes; }

int updown_train_step(const vqa_train_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.B >= 1 && a.K >= 1 && a.K <= 64 && a.V >= 8 && a.H >= 8 && a.A >= 1 && a.T >= 1 && a.E >= 1, "vqa_updown_train_step: bad dims");
  VQA_REQUIRE(a.V % 8 == 0 && a.H % 8 == 0, "vqa_updown_train_step: V and H must be multiples of 8");
  VQA_REQUIRE(2 * a.H + 2 * a.H <= SVEC && a.A <= SVEC && a.V <= SVEC, "vqa_updown_train_step: dims exceed %d", SVEC);
  VQA_REQUIRE(a.dropout_att >= 0.f && a.dropout_att < 1.f && a.dropout_cls >= 0.f && a.dropout_cls < 1.f, "vqa_updown_train_step: dropout");
  VQA_REQUIRE(a.d_img && a.d_tokens && a.d_target && a.d_loss && a.d_logits, "vqa_updown_train_step: NULL input/output");
  VQA_REQUIRE(a.p_emb && a.p_w_ih && a.p_w_hh && a.p_b_ih && a.p_b_hh && a.g_emb && a.g_w_ih && a.g_w_hh && a.g_b_ih && a.g_b_hh,
              "vqa_updown_train_step: NULL GRU parameter / gradient");
  for (int l = 0; l < NL; ++l)
    VQA_REQUIRE(a.p_v[l] && a.p_g[l] && a.p_b[l] && a.g_v[l] && a.g_g[l] && a.g_b[l], "vqa_updown_train_step: NULL layer %d", l);
  const TrainWs need = carve_train(a, nullptr);
  VQA_REQUIRE(a.d_workspace && a.workspace_bytes >= need.bytes, "vqa_updown_train_step: workspace %zu < %zu bytes", a.workspace_bytes,
              need.bytes);
  const TrainWs w = carve_train(a, a.d_workspace);
  if (a.dtype == VQA_BF16) return train_step_t<__nv_bfloat16>(a, w, s);
  if (a.dtype == VQA_F32) return train_step_t<float>(a, w, s);
  return fail(VQA_ERR_INVALID, "vqa_updown_train_step: dtype=%d", a.dtype);
}

}  // namespace vqa
