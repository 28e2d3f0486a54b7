// k6 top-down attention pooling, a14 lowest-index argmax, GRU gate update,
// embedding gather and f32<->bf16 casts.  All HBM-bound streaming kernels:
// 16-byte vector accesses, one CTA per image / row group, no tensor cores.
#include <stdlib.h>

#include "tc_common.cuh"

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

// ---------------------------------------------------------------------------
// k6, streaming form (att + vsum, the Up-Down forward): persistent CTAs; one image's [K,V] block is CONTIGUOUS in
// memory (147 KB at K=36, V=2048 bf16), so it is streamed as a few large sequential bulk async copies
// (cp.async.bulk, mbarrier complete_tx) through a shared-memory ring instead of as K strided 1 KB pieces per CTA
// through registers (the kernel above; ncu cold: 41 % of DRAM peak, 3.4 TB/s).
//   warps 0..7  consumers: thread t owns channels [8t, 8t+8) for the whole image (V <= 2048)
//   warp 8      producer: one lane issues one bulk copy per ring stage (R whole rows, <= 36 KB)
//   warp 0 also computes the image's softmax (double-buffered in shared memory) before the first chunk is consumed
// ---------------------------------------------------------------------------
constexpr int kStreamMaxStages = 6;
constexpr int kStreamChunkBytes = 36 * 1024;
constexpr int kStreamConsumers = 256;
constexpr int kStreamThreads = kStreamConsumers + 32;

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void lds8(const float* p, float (&v)[8]) {
  const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}

template <typename T>
__global__ void __launch_bounds__(kStreamThreads, 1)
attention_pool_stream_kernel(const float* __restrict__ parts, int n_parts, float bias, const T* __restrict__ x, int B, int K,
                             int V, int rev, int S, int R, float* __restrict__ att_out, T* __restrict__ vsum) {
  using namespace tc;
  extern __shared__ __align__(128) uint8_t ring_smem[];
  const uint32_t row_bytes = (uint32_t)V * sizeof(T);
  const uint32_t chunk_bytes = (uint32_t)R * row_bytes;
  const int chunks = (K + R - 1) / R;
  const uint32_t ring = smem_u32(ring_smem);
  const uint32_t bars = ring + (uint32_t)S * chunk_bytes;                   // full[0..S), empty[0..S)
  float* s_att = reinterpret_cast<float*>(ring_smem + (size_t)S * chunk_bytes + 128);   // [2][kPoolMaxK]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int st = 0; st < S; ++st) { mbar_init(bars + 8u * st, 1); mbar_init(bars + 8u * (S + st), kStreamConsumers / 32); }
    fence_barrier_init();
  }
  __syncthreads();
  griddep_launch();
  griddep_wait();
  if (warp == kStreamConsumers / 32) {
    // ===== producer =====
    if (lane == 0) {
      int it = 0;
      for (int img = blockIdx.x; img < B; img += gridDim.x) {
        const int b = rev ? B - 1 - img : img;                    // descending image order, see attention_pool_kernel
        const T* src = x + (size_t)b * K * V;
        for (int c = 0; c < chunks; ++c, ++it) {
          const int st = it % S;
          mbar_wait(bars + 8u * (S + st), (((uint32_t)(it / S)) & 1u) ^ 1u);
          const int rows = (K - c * R) < R ? (K - c * R) : R;
          const uint32_t bytes = (uint32_t)rows * row_bytes;
          mbar_arrive_expect_tx(bars + 8u * st, bytes);
          bulk_copy_g2s(ring + (uint32_t)st * chunk_bytes, src + (size_t)c * R * V, bytes, bars + 8u * st);
        }
      }
    }
  } else {
    // ===== consumers =====
    const int ch = threadIdx.x * 8;
    const bool active = ch < V;
    int it = 0, n = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x, ++n) {
      const int b = rev ? B - 1 - img : img;
      float* att = s_att + (n & 1) * kPoolMaxK;
      if (warp == 0) {
        // softmax over the K regions (lane handles k = lane and lane + 32) while the copies land
        float l0 = -INFINITY, l1 = -INFINITY;
        if (lane < K) {
          const float* p = parts + (size_t)(b * K + lane) * n_parts;
          float sacc = 0.f;
          for (int i = 0; i < n_parts; ++i) sacc += __ldg(p + i);
          l0 = sacc + bias;
        }
        if (lane + 32 < K) {
          const float* p = parts + (size_t)(b * K + lane + 32) * n_parts;
          float sacc = 0.f;
          for (int i = 0; i < n_parts; ++i) sacc += __ldg(p + i);
          l1 = sacc + bias;
        }
        const float m = warp_max(fmaxf(l0, l1));
        const float e0 = (lane < K) ? expf(l0 - m) : 0.f;
        const float e1 = (lane + 32 < K) ? expf(l1 - m) : 0.f;
        const float inv = 1.f / warp_sum(e0 + e1);
        if (lane < K) { att[lane] = e0 * inv; if (att_out) att_out[(size_t)b * K + lane] = e0 * inv; }
        if (lane + 32 < K) { att[lane + 32] = e1 * inv; if (att_out) att_out[(size_t)b * K + lane + 32] = e1 * inv; }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumers) : "memory");
      float acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
      for (int c = 0; c < chunks; ++c, ++it) {
        const int st = it % S;
        mbar_wait(bars + 8u * st, ((uint32_t)(it / S)) & 1u);
        const int rows = (K - c * R) < R ? (K - c * R) : R;
        if (active) {
          const T* base = reinterpret_cast<const T*>(ring_smem + (size_t)st * chunk_bytes) + ch;
#pragma unroll 3
          for (int r = 0; r < rows; ++r) {
            float v[8];
            lds8(base + (size_t)r * V, v);
            const float a = att[c * R + r];
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, v[i], acc[i]);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8u * (S + st));          // this warp is done with the stage
      }
      if (active) store8(vsum + (size_t)b * V + ch, acc);
    }
  }
}

static bool pool_stream_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("VQA_B200_POOL_STREAM"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

template <typename T>
static int launch_pool_stream(const float* parts, int n_parts, float bias, const T* x, int B, int K, int V, int rev, float* att,
                              T* vsum, cudaStream_t s) {
  const int row_bytes = V * (int)sizeof(T);
  int R = kStreamChunkBytes / row_bytes;
  if (R < 1) R = 1;
  if (R > K) R = K;
  const int chunks = (K + R - 1) / R;
  int S = kStreamMaxStages;
  const size_t smem = (size_t)S * R * row_bytes + 128 + 2 * kPoolMaxK * sizeof(float);
  const int grid = B < sm_count() ? B : sm_count();
  (void)chunks;
  static DeviceOnce attr_done[2];                    // per device, not per process
  const int dev = current_device();
  if (attr_done[sizeof(T) == 2].need(dev)) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(attention_pool_stream_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kStreamMaxStages * kStreamChunkBytes + 1024));
    attr_done[sizeof(T) == 2].mark(dev);
  }
  VQA_CUDA_CHECK(launch_pdl(attention_pool_stream_kernel<T>, dim3(grid), dim3(kStreamThreads), smem, s, parts, n_parts, bias,
                            x, B, K, V, rev, S, R, att, vsum));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// fp32-class mode (VQA_F16X2): the features and the pooled result are fp16 plane pairs, x = hi + lo'·2^-11 (exact in
// f32); same work split as attention_pool_kernel, twice the 16-byte loads per thread (as many bytes as f32 features)
__global__ void __launch_bounds__(kPoolThreads)
attention_pool_split_kernel(const float* __restrict__ parts, int n_parts, float bias, const __half* __restrict__ x_hi,
                            const __half* __restrict__ x_lo, int B, int K, int V, int slices, int rev,
                            float* __restrict__ att_out, __half* __restrict__ vsum_hi, __half* __restrict__ vsum_lo) {
  __shared__ float s_att[kPoolMaxK];
  griddep_launch();
  griddep_wait();
  const int b = rev ? B - 1 - (int)(blockIdx.x / slices) : (int)(blockIdx.x / slices);
  const int slice = blockIdx.x % slices;
  const int tid = threadIdx.x;
  if (tid < 32) {
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
  if (vsum_hi == nullptr) return;
  const int c = slice * kPoolChan + tid * 8;
  if (c >= V) return;
  const size_t off = (size_t)b * K * V + c;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 6
  for (int k = 0; k < K; ++k) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(x_hi + off + (size_t)k * V));
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(x_lo + off + (size_t)k * V));
    const __half2* hh = reinterpret_cast<const __half2*>(&h);
    const __half2* ll = reinterpret_cast<const __half2*>(&l);
    const float a = s_att[k];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 fh = __half22float2(hh[i]), fl = __half22float2(ll[i]);
      acc[2 * i] = fmaf(a, fmaf(fl.x, 0x1p-11f, fh.x), acc[2 * i]);
      acc[2 * i + 1] = fmaf(a, fmaf(fl.y, 0x1p-11f, fh.y), acc[2 * i + 1]);
    }
  }
  uint4 oh, ol;
  uint32_t* ph = reinterpret_cast<uint32_t*>(&oh); uint32_t* pl = reinterpret_cast<uint32_t*>(&ol);
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::split_f16x2_pair(acc[2 * i], acc[2 * i + 1], ph[i], pl[i]);
  *reinterpret_cast<uint4*>(vsum_hi + (size_t)b * V + c) = oh;
  *reinterpret_cast<uint4*>(vsum_lo + (size_t)b * V + c) = ol;
}

int attention_pool(const float* parts, int n_parts, float bias, const void* x, int B, int K, int V,
                   int dtype, float* att, void* vsum, void* vatt, cudaStream_t s) {
  VQA_REQUIRE(K >= 1 && K <= kPoolMaxK, "attention_pool: K=%d out of range", K);
  VQA_REQUIRE(V % 8 == 0 && n_parts >= 1, "attention_pool: V=%d must be a multiple of 8", V);
  if (B == 0) return VQA_OK;
  VQA_REQUIRE(parts && x, "attention_pool: NULL input");
  if (dtype == VQA_F16X2) {
    VQA_REQUIRE(vatt == nullptr, "attention_pool(f16x2): the per-region product is not built for plane pairs");
    const int slices = vsum != nullptr ? (V + kPoolChan - 1) / kPoolChan : 1;
    const __half* x_hi = (const __half*)x;
    __half* v_hi = (__half*)vsum;
    VQA_CUDA_CHECK(launch_pdl(attention_pool_split_kernel, dim3((unsigned)B * slices), dim3(kPoolThreads), 0, s, parts, n_parts,
                              bias, x_hi, x_hi + (size_t)B * K * V, B, K, V, slices, l2_order_enabled() ? 1 : 0, att, v_hi,
                              v_hi ? v_hi + (size_t)B * V : nullptr));
    VQA_LAUNCH_CHECK();
    return VQA_OK;
  }
  if (vsum != nullptr && vatt == nullptr && pool_stream_enabled() && V <= kStreamConsumers * 8 &&
      V * (int)elem_size(dtype) <= kStreamChunkBytes && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(vsum) & 15) == 0) {
    const int rev = l2_order_enabled() ? 1 : 0;
    if (dtype == VQA_BF16)
      return launch_pool_stream<__nv_bfloat16>(parts, n_parts, bias, (const __nv_bfloat16*)x, B, K, V, rev, att,
                                               (__nv_bfloat16*)vsum, s);
    return launch_pool_stream<float>(parts, n_parts, bias, (const float*)x, B, K, V, rev, att, (float*)vsum, s);
  }
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
                        const uint4* __restrict__ emb, uint4* __restrict__ out, uint4* __restrict__ zero, int zero_vec16) {
  // first kernel of the forward chain: it also clears the small must-be-zero scratch parked behind the GRU workspace
  // (keys / counters of the fused answer selection) instead of a memset node of its own
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_vec16; i += gridDim.x * blockDim.x) zero[i] = make_uint4(0, 0, 0, 0);
  const int total = n_rows * row_vec16;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / row_vec16, c = i - r * row_vec16;
    long long tok = tokens[r];
    tok = tok < 0 ? 0 : (tok >= ntoken_rows ? ntoken_rows - 1 : tok);
    out[i] = __ldg(emb + (size_t)tok * row_vec16 + c);
  }
}

int embedding_gather(const int64_t* tokens, int n_rows, int E_pad, int ntoken_rows, int dtype,
                     const void* emb, void* out, cudaStream_t s, void* zero_ptr, size_t zero_bytes) {
  const size_t row_bytes = (size_t)E_pad * elem_size(dtype);
  VQA_REQUIRE(row_bytes % 16 == 0, "embedding_gather: padded row of %zu bytes is not 16-byte aligned",
              row_bytes);
  if (n_rows == 0) return VQA_OK;
  const int row_vec16 = (int)(row_bytes / 16);
  const int total = n_rows * row_vec16;
  int grid = (total + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  VQA_REQUIRE(zero_bytes % 16 == 0 && ((uintptr_t)zero_ptr & 15) == 0, "embedding_gather: zero region must be 16-byte aligned");
  embedding_gather_kernel<<<grid, 256, 0, s>>>(tokens, n_rows, row_vec16, ntoken_rows,
                                               (const uint4*)emb, (uint4*)out, (uint4*)zero_ptr, (int)(zero_bytes / 16));
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

// fp32-class mode: gi row = gi_table[token(b, t)] (f32 [rows, 3H] = W_ih·emb[v] + b_ih), gh = this step's recurrent GEMM
// (NULL at t = 0: h_0 = 0, gh = b_hh); the new state leaves as f32 and as the fp16 plane pair that feeds the next GEMM
__global__ void __launch_bounds__(256)
gru_gate_table_kernel(const float* __restrict__ gi_table, const int64_t* __restrict__ tokens, int ntoken_rows,
                      const float* __restrict__ gh, const float* __restrict__ b_hh, int B, int H, int Tlen, int t,
                      const float* h_prev, float* h_out, __half* __restrict__ h_hi, __half* __restrict__ h_lo) {
  griddep_launch();
  griddep_wait();
  const int total = B * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / H, j = i - b * H;
    long long tok = tokens[(size_t)b * Tlen + t];
    tok = tok < 0 ? 0 : (tok >= ntoken_rows ? ntoken_rows - 1 : tok);
    const float* gir = gi_table + (size_t)tok * 3 * H;
    const float* ghr = gh ? gh + (size_t)b * 3 * H : b_hh;
    const float r = 1.f / (1.f + expf(-(gir[j] + ghr[j])));
    const float z = 1.f / (1.f + expf(-(gir[H + j] + ghr[H + j])));
    const float n = tanhf(gir[2 * H + j] + r * ghr[2 * H + j]);
    const float hn = (1.f - z) * n + z * (h_prev ? h_prev[i] : 0.f);
    h_out[i] = hn;
    const __half hh = __float2half_rn(hn);
    h_hi[i] = hh;
    h_lo[i] = __float2half_rn((hn - __half2float(hh)) * 2048.f);
  }
}

int gru_gate_table(const float* gi_table, const int64_t* tokens, int ntoken_rows, const float* gh, const float* b_hh, int B,
                   int H, int T, int t, const float* h_prev, float* h_out, void* h_planes, cudaStream_t s) {
  const int total = B * H;
  int grid = (total + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  __half* hi = (__half*)h_planes;
  VQA_CUDA_CHECK(launch_pdl(gru_gate_table_kernel, dim3(grid), dim3(256), 0, s, gi_table, tokens, ntoken_rows, gh, b_hh, B, H, T,
                            t, h_prev, h_out, hi, hi + (size_t)B * H));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
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
// LSTM gate update (torch LSTM semantics, modules.py:123-130 with rnn_type='LSTM'), gate order [i; f; g; o]:
//   c' = σ(f) ⊙ c + σ(i) ⊙ tanh(g),  h' = σ(o) ⊙ tanh(c')     (gates = W_ih x + b_ih + W_hh h + b_hh, all four summed
//   before this kernel: the input half enters the recurrent GEMM as its additive epilogue operand)
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
lstm_gate_kernel(const float* __restrict__ gates, int B, int H, float* c, float* h_out, T* __restrict__ h_lp, int ld_lp) {
  const int total = B * H;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int b = i / H, j = i - b * H;
    const float* g = gates + (size_t)b * 4 * H;
    const float ig = 1.f / (1.f + expf(-g[j]));
    const float fg = 1.f / (1.f + expf(-g[H + j]));
    const float gg = tanhf(g[2 * H + j]);
    const float og = 1.f / (1.f + expf(-g[3 * H + j]));
    const float cn = fg * c[i] + ig * gg;
    const float hn = og * tanhf(cn);
    c[i] = cn;
    if (h_out) h_out[i] = hn;
    h_lp[(size_t)b * ld_lp + j] = Elem<T>::from_f(hn);
  }
}

int lstm_gate(const float* gates, int B, int H, float* c, float* h_out, void* h_lp, int ld_lp, int dtype, cudaStream_t s) {
  const int total = B * H;
  int grid = (total + 255) / 256;
  if (grid > sm_count() * 8) grid = sm_count() * 8;
  if (dtype == VQA_BF16)
    lstm_gate_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(gates, B, H, c, h_out, (__nv_bfloat16*)h_lp, ld_lp);
  else
    lstm_gate_kernel<float><<<grid, 256, 0, s>>>(gates, B, H, c, h_out, (float*)h_lp, ld_lp);
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

// f32 -> fp16 plane pair (VQA_F16X2): hi = fp16(x), lo' = fp16((x - hi)·2^11); n contiguous elements, streaming
__global__ void __launch_bounds__(256)
split_f32_kernel(const float* __restrict__ src, __half* __restrict__ hi, __half* __restrict__ lo, size_t n) {
  const size_t n8 = n / 8;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldcs(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    const float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint4 h, l;
    uint32_t* ph = reinterpret_cast<uint32_t*>(&h); uint32_t* pl = reinterpret_cast<uint32_t*>(&l);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn((x[2 * e] - hf.x) * 2048.f, (x[2 * e + 1] - hf.y) * 2048.f);
      ph[e] = *reinterpret_cast<const uint32_t*>(&hh);
      pl[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    reinterpret_cast<uint4*>(hi)[i] = h;
    reinterpret_cast<uint4*>(lo)[i] = l;
  }
  if (blockIdx.x == 0)
    for (size_t i = n8 * 8 + threadIdx.x; i < n; i += blockDim.x) {
      const __half h = __float2half_rn(src[i]);
      hi[i] = h; lo[i] = __float2half_rn((src[i] - __half2float(h)) * 2048.f);
    }
}

int split_f32(const float* src, void* hi, void* lo, size_t n, cudaStream_t s) {
  VQA_REQUIRE(((uintptr_t)src % 16 == 0) && ((uintptr_t)hi % 16 == 0) && ((uintptr_t)lo % 16 == 0), "split: pointers must be 16-byte aligned");
  if (n == 0) return VQA_OK;
  size_t grid = (n / 8 + 255) / 256;
  if (grid < 1) grid = 1;
  if (grid > (size_t)sm_count() * 16) grid = (size_t)sm_count() * 16;
  split_f32_kernel<<<(unsigned)grid, 256, 0, s>>>(src, (__half*)hi, (__half*)lo, n);
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
