// FFMA tile implementation of vqa_linear — the fp32 mode (VQA_F32) of the
// weight-normed linear layer (modules.py:13-60) and the bring-up / cross-check
// implementation for the tcgen05 kernel (VQA_B200_FORCE_SIMT=1 routes bf16
// operands here).  128x128x16 tiles, 256 threads, 8x8 outputs per thread, fp32
// accumulate, same fused epilogue contract as the tensor-core kernel.
#include "common.cuh"

namespace vqa {

constexpr int SBM = 128, SBN = 128, SBK = 16, STHREADS = 256;

struct Epilogue {
  const float* scale; const float* bias; int relu;
  const float* mul; int ld_mul; int mul_row_div;
  const float* add; int ld_add; int add_row_div;
  const void* mask; int ld_mask; int mask_bf16;
  const float* logit_w;
  void* out; int ldo; int out_dtype; int n_parts;
  float leaky_slope; int add_after_act; int sigmoid;
};

// TA / TW: the operand is given with the contraction index as the ROW index ([K,M] / [K,N]
// row-major) — the backward GEMMs of the training step; guarded scalar loads (parity path only).
template <typename T, bool TA, bool TW>
__global__ void __launch_bounds__(STHREADS)
linear_simt_kernel(const T* __restrict__ A, int lda, const T* __restrict__ W, int ldw, int M, int N,
                   int K, Epilogue ep) {
  __shared__ float As[SBK][SBM + 4];
  __shared__ float Ws[SBK][SBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * SBM, n0 = blockIdx.x * SBN;
  const int tm = tid / 16, tn = tid % 16;          // 16x16 threads, 8x8 outputs each
  const int lrow = tid / 2, lk = (tid % 2) * 8;    // loader: 128 rows x 2 chunks of 8

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int arow = m0 + lrow, wrow = n0 + lrow;
  const bool a_ok = arow < M, w_ok = wrow < N;
  const T* ap = A + (size_t)(a_ok ? arow : 0) * lda + lk;
  const T* wp = W + (size_t)(w_ok ? wrow : 0) * ldw + lk;

  const int tk = tid / 16, tc8 = (tid % 16) * 8;  // transposed loader: 16 k-rows x 16 chunks of 8 columns
  for (int k0 = 0; k0 < K; k0 += SBK) {
    if constexpr (!TA) {
      float av[8];
      if (k0 + lk + 8 <= K) load8(ap + k0, av);
      else
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = (k0 + lk + i < K) ? Elem<T>::to_f(ap[k0 + i]) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) As[lk + i][lrow] = a_ok ? av[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = m0 + tc8 + i, k = k0 + tk;
        As[tk][tc8 + i] = (m < M && k < K) ? Elem<T>::to_f(A[(size_t)k * lda + m]) : 0.f;
      }
    }
    if constexpr (!TW) {
      float wv[8];
      if (k0 + lk + 8 <= K) load8(wp + k0, wv);
      else
#pragma unroll
        for (int i = 0; i < 8; ++i) wv[i] = (k0 + lk + i < K) ? Elem<T>::to_f(wp[k0 + i]) : 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) Ws[lk + i][lrow] = w_ok ? wv[i] : 0.f;
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = n0 + tc8 + i, k = k0 + tk;
        Ws[tk][tc8 + i] = (n < N && k < K) ? Elem<T>::to_f(W[(size_t)k * ldw + n]) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SBK; ++k) {
      float a[8], w[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][tm * 8 + i];
#pragma unroll
      for (int j = 0; j < 8; ++j) w[j] = Ws[k][tn * 8 + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }

  // fused epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + tm * 8 + i;
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + tn * 8 + j;
      if (m < M && n < N) {
        float y = acc[i][j];
        if (ep.scale) y *= ep.scale[n];
        if (ep.bias) y += ep.bias[n];
        if (ep.add && !ep.add_after_act) y += ep.add[(size_t)(m / ep.add_row_div) * ep.ld_add + n];
        if (ep.relu) y = ep.leaky_slope == 0.f ? fmaxf(y, 0.f) : (y > 0.f ? y : ep.leaky_slope * y);
        if (ep.add && ep.add_after_act) y += ep.add[(size_t)(m / ep.add_row_div) * ep.ld_add + n];
        if (ep.mul) y *= ep.mul[(size_t)(m / ep.mul_row_div) * ep.ld_mul + n];
        if (ep.mask) {
          const float mk = ep.mask_bf16 ? __bfloat162float(((const __nv_bfloat16*)ep.mask)[(size_t)m * ep.ld_mask + n])
                                        : ((const float*)ep.mask)[(size_t)m * ep.ld_mask + n];
          if (!(mk > 0.f)) y = 0.f;
        }
        if (ep.sigmoid) y = 1.f / (1.f + expf(-y));
        if (ep.logit_w) {
          part = fmaf(y, ep.logit_w[n], part);
        } else if (ep.out_dtype == VQA_BF16) {
          ((__nv_bfloat16*)ep.out)[(size_t)m * ep.ldo + n] = __float2bfloat16_rn(y);
        } else {
          ((float*)ep.out)[(size_t)m * ep.ldo + n] = y;
        }
      }
    }
    if (ep.logit_w) {
      // the 16 threads of one tm are 16 consecutive lanes: reduce within the half warp
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (tn == 0 && m < M) ((float*)ep.out)[(size_t)m * ep.n_parts + blockIdx.x] = part;
    }
  }
}

template <typename T>
static void launch_simt(const vqa_linear_args& a, const Epilogue& ep, dim3 grid, cudaStream_t s) {
  const T* A = (const T*)a.d_A;
  const T* W = (const T*)a.d_W;
  if (a.trans_a && a.trans_w) linear_simt_kernel<T, true, true><<<grid, STHREADS, 0, s>>>(A, a.lda, W, a.ldw, a.M, a.N, a.K, ep);
  else if (a.trans_w) linear_simt_kernel<T, false, true><<<grid, STHREADS, 0, s>>>(A, a.lda, W, a.ldw, a.M, a.N, a.K, ep);
  else linear_simt_kernel<T, false, false><<<grid, STHREADS, 0, s>>>(A, a.lda, W, a.ldw, a.M, a.N, a.K, ep);
}

int linear_simt(const vqa_linear_args& a, cudaStream_t s) {
  const int esz = (int)elem_size(a.dtype);
  VQA_REQUIRE((a.lda * esz) % 16 == 0 && (a.ldw * esz) % 16 == 0 &&
                  (uintptr_t)a.d_A % 16 == 0 && (uintptr_t)a.d_W % 16 == 0,
              "vqa_linear(simt): A/W rows must be 16-byte aligned (lda=%d ldw=%d)", a.lda, a.ldw);
  VQA_REQUIRE(!(a.trans_a && !a.trans_w), "vqa_linear(simt): trans_a without trans_w is not built");
  if (a.M == 0 || a.N == 0) return VQA_OK;
  Epilogue ep{a.d_scale, a.d_bias, a.relu, a.d_mul, a.ld_mul, a.mul_row_div > 0 ? a.mul_row_div : 1,
              a.d_add, a.ld_add, a.add_row_div > 0 ? a.add_row_div : 1,
              a.d_mask, a.ld_mask, a.mask_dtype == VQA_BF16, a.d_logit_w, a.d_out, a.ldo, a.out_dtype,
              (a.N + SBN - 1) / SBN, a.leaky_slope, a.add_after_act, a.sigmoid};
  dim3 grid((a.N + SBN - 1) / SBN, (a.M + SBM - 1) / SBM);
  if (a.dtype == VQA_BF16) launch_simt<__nv_bfloat16>(a, ep, grid, s);
  else launch_simt<float>(a, ep, grid, s);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
