// k9+k10 — relation-masked K×K graph attention for one CorrelatedGraphConv
// layer, consuming the wide projection Y = x·[W0+W1 ; W2 ; Wa ; Wb]ᵀ.
//
// Reference: gcn.py:93-107 (DirectedGraphConv.conv incl. the label-bias gather),
// modules.py:86-95 (DotProduct), gcn.py:119-128 (relation_alpha: ReLU, adj·α,
// softmax over dim=1 = the ROW index), gcn.py:152-168 (forward), gcn.py:211-212
// (ReLU after the layer) and predictor.py:85 (Σ_K) when vsum is requested.
//
// With f_i = a_i·x_i (a = top-down attention; row scaling commutes with the
// bias-free maps) and P=(W0+W1)x, S=W2x, A'=Wa x, B'=Wb x read from Y:
//   dot_ij  = a_i a_j (A'_i·B'_j) + a_i (A'_i·bb) + a_j (ba·B'_j) + ba·bb
//   α       = softmax_i( Σ_k adj_ik · ReLU(dot_kj) )
//   conv_j  = a_j S_j + Σ_{k: adj_jk} a_k P_k + Σ_l hist_jl · bias_l
//   out_i   = ReLU( Σ_j α_ij conv_j ),   vsum = Σ_i out_i
// hist_jl = #{k : label_jk = l} replaces the reference's [B,K,K,V] gather
// (label 0, incl. the diagonal, contributes bias_0 — gcn.py:107).
//
// One CTA per image, everything K×K stays in shared memory; adjacency rows are
// 64-bit masks so neighbour sums are predicated register adds.  This is the
// FFMA version (v1); the tcgen05 version of the three K×K×V contractions is
// the next step for this kernel.
#include "common.cuh"

namespace vqa {

constexpr int GK = 36;            // regions per image (compile-time: register tiles)
constexpr int GTHREADS = 256;
constexpr int GCH1 = 64;          // d-chunk of phase 1
constexpr int GCH3 = 256;         // channel chunk of phase 3 (= threads)
constexpr int GMAXL = 16;

struct GraphSmem {
  float a[GK];                    // attention scalars
  float ua[GK], ub[GK];           // A'_i·bb, ba·B'_j
  float G[GK][GK + 1];            // A'·B'ᵀ → α0 → α
  float alpha[GK][GK + 4];        // α (padded for float4 reads)
  float hist[GK][GMAXL];
  unsigned long long adj[GK];
  float red[8];
  union {
    struct { float A[GK][GCH1 + 1]; float B[GK][GCH1 + 1]; } p1;
    struct { float conv[GK][GCH3]; } p3;
  } u;
};

template <typename T>
__global__ void __launch_bounds__(GTHREADS)
graph_attention_kernel(const T* __restrict__ Y, int ldy, const float* __restrict__ att,
                       const uint8_t* __restrict__ labels, const float* __restrict__ label_bias,
                       int num_labels, const float* __restrict__ ba, const float* __restrict__ bb,
                       int B, int V, T* __restrict__ out, T* __restrict__ vsum,
                       float* __restrict__ alpha_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GraphSmem& sm = *reinterpret_cast<GraphSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const int b = blockIdx.x;
  const T* Yb = Y + (size_t)b * GK * ldy;
  const T* Pm = Yb;               // (W0+W1) x
  const T* Sm = Yb + V;           // W2 x
  const T* Am = Yb + 2 * V;       // Wa x
  const T* Bm = Yb + 3 * V;       // Wb x

  // ---- phase 0: scalars, adjacency masks, label histogram -------------------
  if (tid < GK) {
    sm.a[tid] = att ? att[(size_t)b * GK + tid] : 1.f;
    const uint8_t* lr = labels + ((size_t)b * GK + tid) * GK;
    unsigned long long m = 0ull;
    float h[GMAXL];
#pragma unroll
    for (int l = 0; l < GMAXL; ++l) h[l] = 0.f;
    for (int k = 0; k < GK; ++k) {
      const int l = lr[k];
      if (l != 0) m |= (1ull << k);
#pragma unroll
      for (int q = 0; q < GMAXL; ++q) h[q] += (q == l) ? 1.f : 0.f;
    }
    sm.adj[tid] = m;
#pragma unroll
    for (int l = 0; l < GMAXL; ++l) sm.hist[tid][l] = h[l];
  }
  // c0 = ba·bb (block reduction)
  float c0 = 0.f;
  for (int d = tid; d < V; d += GTHREADS) c0 = fmaf(ba[d], bb[d], c0);
  c0 = warp_sum(c0);
  if ((tid & 31) == 0) sm.red[tid >> 5] = c0;

  // ---- phase 1: G = A'·B'ᵀ, ua = A'·bb, ub = B'·ba over d-chunks -------------
  // threads 0..143: 3x3 blocks of G; 144..179: ua_i; 180..215: ub_j
  float g[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) g[i][j] = 0.f;
  float uacc = 0.f;
  const int bi = (tid / 12) * 3, bj = (tid % 12) * 3;
  for (int d0 = 0; d0 < V; d0 += GCH1) {
    __syncthreads();
    for (int v = tid; v < 2 * GK * (GCH1 / 8); v += GTHREADS) {
      const int which = v / (GK * (GCH1 / 8));
      const int r = (v / (GCH1 / 8)) % GK, c = (v % (GCH1 / 8)) * 8;
      float x[8];
      load8((which ? Bm : Am) + (size_t)r * ldy + d0 + c, x);
      float* dst = which ? &sm.u.p1.B[r][c] : &sm.u.p1.A[r][c];
#pragma unroll
      for (int i = 0; i < 8; ++i) dst[i] = x[i];
    }
    __syncthreads();
    if (tid < 144) {
#pragma unroll 8
      for (int d = 0; d < GCH1; ++d) {
        float av[3], bv[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) av[i] = sm.u.p1.A[bi + i][d];
#pragma unroll
        for (int j = 0; j < 3; ++j) bv[j] = sm.u.p1.B[bj + j][d];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) g[i][j] = fmaf(av[i], bv[j], g[i][j]);
      }
    } else if (tid < 144 + GK) {
      const int i = tid - 144;
      for (int d = 0; d < GCH1; ++d) uacc = fmaf(sm.u.p1.A[i][d], bb[d0 + d], uacc);
    } else if (tid < 144 + 2 * GK) {
      const int j = tid - 144 - GK;
      for (int d = 0; d < GCH1; ++d) uacc = fmaf(sm.u.p1.B[j][d], ba[d0 + d], uacc);
    }
  }
  if (tid < 144) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) sm.G[bi + i][bj + j] = g[i][j];
  } else if (tid < 144 + GK) {
    sm.ua[tid - 144] = uacc;
  } else if (tid < 144 + 2 * GK) {
    sm.ub[tid - 144 - GK] = uacc;
  }
  __syncthreads();

  // ---- phase 2: α0 = ReLU(dot); α1 = adj·α0; α = softmax over rows i ---------
  c0 = 0.f;
#pragma unroll
  for (int w = 0; w < GTHREADS / 32; ++w) c0 += sm.red[w];
  for (int t = tid; t < GK * GK; t += GTHREADS) {
    const int i = t / GK, j = t - i * GK;
    const float ai = sm.a[i], aj = sm.a[j];
    const float dot = ai * aj * sm.G[i][j] + ai * sm.ua[i] + aj * sm.ub[j] + c0;
    sm.alpha[i][j] = fmaxf(dot, 0.f);              // α0 (temporarily in alpha)
  }
  __syncthreads();
  for (int t = tid; t < GK * GK; t += GTHREADS) {
    const int i = t / GK, j = t - i * GK;
    const unsigned long long m = sm.adj[i];
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < GK; ++k) s += ((m >> k) & 1ull) ? sm.alpha[k][j] : 0.f;
    sm.G[i][j] = s;                                // α1
  }
  __syncthreads();
  if (tid < GK) {                                  // column j = tid: softmax over i
    const int j = tid;
    float mx = -INFINITY;
    for (int i = 0; i < GK; ++i) mx = fmaxf(mx, sm.G[i][j]);
    float sum = 0.f;
    for (int i = 0; i < GK; ++i) { const float e = expf(sm.G[i][j] - mx); sm.G[i][j] = e; sum += e; }
    const float inv = 1.f / sum;
    for (int i = 0; i < GK; ++i) sm.alpha[i][j] = sm.G[i][j] * inv;
  }
  __syncthreads();
  if (tid < GK) {
#pragma unroll
    for (int j = GK; j < GK + 4; ++j) sm.alpha[tid][j] = 0.f;
  }
  if (alpha_out != nullptr)
    for (int t = tid; t < GK * GK; t += GTHREADS)
      alpha_out[(size_t)b * GK * GK + t] = sm.alpha[t / GK][t % GK];

  // ---- phase 3: conv and out, 256 channels at a time (thread = channel) ------
  for (int c0ch = 0; c0ch < V; c0ch += GCH3) {
    const int c = c0ch + tid;
    __syncthreads();
    float p[GK];
#pragma unroll
    for (int k = 0; k < GK; ++k) p[k] = sm.a[k] * Elem<T>::to_f(Pm[(size_t)k * ldy + c]);
    float lb[GMAXL];
#pragma unroll
    for (int l = 0; l < GMAXL; ++l) lb[l] = (l < num_labels) ? label_bias[(size_t)l * V + c] : 0.f;
    for (int j = 0; j < GK; ++j) {
      const unsigned long long m = sm.adj[j];
      float s = sm.a[j] * Elem<T>::to_f(Sm[(size_t)j * ldy + c]);
#pragma unroll
      for (int k = 0; k < GK; ++k) s += ((m >> k) & 1ull) ? p[k] : 0.f;
      const float4* hj = reinterpret_cast<const float4*>(sm.hist[j]);
#pragma unroll
      for (int l4 = 0; l4 < GMAXL / 4; ++l4) {
        const float4 h = hj[l4];
        s = fmaf(h.x, lb[4 * l4], s); s = fmaf(h.y, lb[4 * l4 + 1], s);
        s = fmaf(h.z, lb[4 * l4 + 2], s); s = fmaf(h.w, lb[4 * l4 + 3], s);
      }
      sm.u.p3.conv[j][tid] = s;
    }
    // each thread re-reads only its own column of conv: no barrier needed
#pragma unroll
    for (int k = 0; k < GK; ++k) p[k] = sm.u.p3.conv[k][tid];
    float vs = 0.f;
    for (int i = 0; i < GK; ++i) {
      const float4* ar = reinterpret_cast<const float4*>(sm.alpha[i]);
      float s = 0.f;
#pragma unroll
      for (int j4 = 0; j4 < GK / 4; ++j4) {
        const float4 al = ar[j4];
        s = fmaf(al.x, p[4 * j4], s); s = fmaf(al.y, p[4 * j4 + 1], s);
        s = fmaf(al.z, p[4 * j4 + 2], s); s = fmaf(al.w, p[4 * j4 + 3], s);
      }
      s = fmaxf(s, 0.f);
      if (out != nullptr) out[((size_t)b * GK + i) * V + c] = Elem<T>::from_f(s);
      vs += s;
    }
    if (vsum != nullptr) vsum[(size_t)b * V + c] = Elem<T>::from_f(vs);
  }
}

int graph_attention(const vqa_graph_attention_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.d_Y && a.d_labels && a.d_label_bias && a.d_ba && a.d_bb, "graph_attention: NULL input");
  if (a.K != GK) return fail(VQA_ERR_UNSUPPORTED, "graph_attention: K=%d (only K=%d is built)", a.K, GK);
  VQA_REQUIRE(a.V % GCH3 == 0, "graph_attention: V=%d must be a multiple of %d", a.V, GCH3);
  VQA_REQUIRE(a.num_labels >= 1 && a.num_labels <= GMAXL, "graph_attention: num_labels=%d", a.num_labels);
  VQA_REQUIRE(a.ldy >= 4 * a.V && a.ldy % 8 == 0, "graph_attention: ldy=%d", a.ldy);
  if (a.B == 0) return VQA_OK;
  const size_t smem = sizeof(GraphSmem);
  if (a.dtype == VQA_BF16) {
    auto kern = graph_attention_kernel<__nv_bfloat16>;
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, GTHREADS, smem, s>>>((const __nv_bfloat16*)a.d_Y, a.ldy, a.d_att, a.d_labels,
                                     a.d_label_bias, a.num_labels, a.d_ba, a.d_bb, a.B, a.V,
                                     (__nv_bfloat16*)a.d_out, (__nv_bfloat16*)a.d_vsum, a.d_alpha);
  } else {
    auto kern = graph_attention_kernel<float>;
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<a.B, GTHREADS, smem, s>>>((const float*)a.d_Y, a.ldy, a.d_att, a.d_labels, a.d_label_bias,
                                     a.num_labels, a.d_ba, a.d_bb, a.B, a.V, (float*)a.d_out,
                                     (float*)a.d_vsum, a.d_alpha);
  }
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
