// Optimizer side of the training step (BASELINE config 4; the step AFTER the path, train.py:109-111):
//   nn.utils.clip_grad_norm_(model.parameters(), max_norm)  ->  grad_norm_sq + clip_scale kernels
//   torch.optim.Adamax(...).step()                           ->  adamax kernel
// The reference runs these as ~30 small torch launches per step over 26 parameter tensors (≈ 1.5 ms of host-bound
// launches next to a 1.6 ms forward + backward).  Here every parameter tensor of the model is one entry of a
// pointer table passed by value in the kernel parameters ("multi-tensor apply"): ONE launch per operation,
// HBM-streaming with 16-byte accesses (Adamax: 4 reads + 3 writes of 4 B per parameter = 28 B/parameter).
//
// Adamax update (torch.optim.Adamax, single-tensor form, maximize=False):
//   g   = grad * grad_scale (+ weight_decay * p)
//   m  += (1 - beta1) * (g - m)                    (exp_avg.lerp_(g, 1 - beta1))
//   u   = max(beta2 * u, |g| + eps)                (exp_inf)
//   p  -= lr / (1 - beta1^step) * m / u
#include "common.cuh"

namespace vqa {

constexpr int kOptThreads = 256;
constexpr int kOptChunk = kOptThreads * 16;          // elements per block visit

struct OptTable {
  float* p[VQA_OPTIM_MAX_TENSORS];
  const float* g[VQA_OPTIM_MAX_TENSORS];
  float* m[VQA_OPTIM_MAX_TENSORS];
  float* u[VQA_OPTIM_MAX_TENSORS];
  unsigned long long n[VQA_OPTIM_MAX_TENSORS];
  float lr[VQA_OPTIM_MAX_TENSORS];                   // per tensor (param groups, train.py:54-56)
  int chunk_end[VQA_OPTIM_MAX_TENSORS];              // prefix sum of chunk counts
  int count;
};

__device__ __forceinline__ int find_tensor(const OptTable& t, int chunk) {
  int i = 0;
  while (i < t.count - 1 && chunk >= t.chunk_end[i]) ++i;
  return i;
}

// partial[b] = Σ g² over block b's chunks (fixed order: deterministic), finalised by clip_scale_kernel
__global__ void __launch_bounds__(kOptThreads)
grad_norm_sq_kernel(const __grid_constant__ OptTable t, int total_chunks, float* __restrict__ partial) {
  __shared__ float red[kOptThreads / 32];
  float acc = 0.f;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(t, chunk);
    const size_t base = (size_t)(chunk - (ti ? t.chunk_end[ti - 1] : 0)) * kOptChunk;
    const size_t n = t.n[ti];
    const float* g = t.g[ti];
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    for (int k = 0; k < 4; ++k) {
      const size_t i = base + ((size_t)k * kOptThreads + threadIdx.x) * 4;
      if (vec && i + 4 <= n) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + i));
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      } else {
        for (size_t j = i; j < n && j < i + 4; ++j) acc += g[j] * g[j];
      }
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < kOptThreads / 32; ++w) s += red[w];
    partial[blockIdx.x] = s;
  }
}

// total_norm = sqrt(Σ partial); scale = min(1, max_norm / (total_norm + 1e-6))   (clip_grad_norm_'s clamp)
__global__ void __launch_bounds__(32)
clip_scale_kernel(const float* __restrict__ partial, int n_partial, float max_norm, float* __restrict__ total_norm,
                  float* __restrict__ scale) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n_partial; i += 32) s += partial[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) {
    const float nrm = sqrtf(s);
    *total_norm = nrm;
    *scale = fminf(max_norm / (nrm + 1e-6f), 1.f);
  }
}

// in-place g *= scale (what clip_grad_norm_ leaves in .grad); skipped by callers that hand `scale` to adamax instead
__global__ void __launch_bounds__(kOptThreads)
grad_scale_kernel(const __grid_constant__ OptTable t, int total_chunks, const float* __restrict__ scale) {
  const float s = *scale;
  if (s == 1.f) return;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(t, chunk);
    const size_t base = (size_t)(chunk - (ti ? t.chunk_end[ti - 1] : 0)) * kOptChunk;
    const size_t n = t.n[ti];
    float* g = const_cast<float*>(t.g[ti]);
    const bool vec = (reinterpret_cast<uintptr_t>(g) & 15) == 0;
    for (int k = 0; k < 4; ++k) {
      const size_t i = base + ((size_t)k * kOptThreads + threadIdx.x) * 4;
      if (vec && i + 4 <= n) {
        float4 v = *reinterpret_cast<float4*>(g + i);
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        *reinterpret_cast<float4*>(g + i) = v;
      } else {
        for (size_t j = i; j < n && j < i + 4; ++j) g[j] *= s;
      }
    }
  }
}

__device__ __forceinline__ void adamax_one(float& p, float g, float& m, float& u, float gs, float wd, float w1, float beta2,
                                           float eps, float clr) {
  g *= gs;
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = m + w1 * (g - m);
  u = fmaxf(beta2 * u, fabsf(g) + eps);
  p = p - clr * (m / u);
}

__global__ void __launch_bounds__(kOptThreads)
adamax_kernel(const __grid_constant__ OptTable t, int total_chunks, float beta1, float beta2, float eps, float wd,
              const float* __restrict__ grad_scale) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  const float w1 = 1.f - beta1;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    const int ti = find_tensor(t, chunk);
    const size_t base = (size_t)(chunk - (ti ? t.chunk_end[ti - 1] : 0)) * kOptChunk;
    const size_t n = t.n[ti];
    float* p = t.p[ti]; const float* g = t.g[ti]; float* m = t.m[ti]; float* u = t.u[ti];
    const float clr = t.lr[ti];                       // lr / (1 - beta1^step), rounded once on the host like torch's clr
    const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                       reinterpret_cast<uintptr_t>(u)) & 15) == 0;
    for (int k = 0; k < 4; ++k) {
      const size_t i = base + ((size_t)k * kOptThreads + threadIdx.x) * 4;
      if (vec && i + 4 <= n) {
        float4 pv = *reinterpret_cast<float4*>(p + i), mv = *reinterpret_cast<float4*>(m + i), uv = *reinterpret_cast<float4*>(u + i);
        const float4 gv = __ldg(reinterpret_cast<const float4*>(g + i));
        adamax_one(pv.x, gv.x, mv.x, uv.x, gs, wd, w1, beta2, eps, clr);
        adamax_one(pv.y, gv.y, mv.y, uv.y, gs, wd, w1, beta2, eps, clr);
        adamax_one(pv.z, gv.z, mv.z, uv.z, gs, wd, w1, beta2, eps, clr);
        adamax_one(pv.w, gv.w, mv.w, uv.w, gs, wd, w1, beta2, eps, clr);
        *reinterpret_cast<float4*>(p + i) = pv; *reinterpret_cast<float4*>(m + i) = mv; *reinterpret_cast<float4*>(u + i) = uv;
      } else {
        for (size_t j = i; j < n && j < i + 4; ++j) adamax_one(p[j], g[j], m[j], u[j], gs, wd, w1, beta2, eps, clr);
      }
    }
  }
}

static int build_table(const vqa_optim_tensor* ts, int count, bool need_state, OptTable& t, int& total_chunks) {
  VQA_REQUIRE(ts && count >= 1 && count <= VQA_OPTIM_MAX_TENSORS, "optim: tensor count %d outside [1,%d]", count,
              VQA_OPTIM_MAX_TENSORS);
  int chunks = 0;
  for (int i = 0; i < count; ++i) {
    VQA_REQUIRE(ts[i].d_g && (!need_state || (ts[i].d_p && ts[i].d_m && ts[i].d_u)), "optim: NULL pointer in tensor %d", i);
    t.p[i] = ts[i].d_p; t.g[i] = ts[i].d_g; t.m[i] = ts[i].d_m; t.u[i] = ts[i].d_u; t.n[i] = ts[i].n; t.lr[i] = ts[i].lr;
    chunks += (int)((ts[i].n + kOptChunk - 1) / kOptChunk);
    t.chunk_end[i] = chunks;
  }
  t.count = count;
  total_chunks = chunks;
  return VQA_OK;
}

static int grid_for_chunks(int chunks) {
  const int cap = sm_count() * 8;
  return chunks < cap ? (chunks > 0 ? chunks : 1) : cap;
}

size_t optim_norm_workspace_bytes() { return (size_t)(sm_count() * 8) * sizeof(float); }

int grad_clip(const vqa_optim_tensor* ts, int count, float max_norm, int scale_in_place, float* ws, float* total_norm,
              float* scale, cudaStream_t s) {
  VQA_REQUIRE(ws && total_norm && scale, "grad_clip: NULL pointer");
  int done = 0;
  // more tensors than one table holds: partial sums of every group land in consecutive workspace slots
  const int cap = sm_count() * 8;
  int groups = (count + VQA_OPTIM_MAX_TENSORS - 1) / VQA_OPTIM_MAX_TENSORS;
  VQA_REQUIRE(groups >= 1 && groups <= 8, "grad_clip: %d tensors (at most %d)", count, 8 * VQA_OPTIM_MAX_TENSORS);
  const int per_group = cap / groups;
  int used = 0;
  for (int gi = 0; gi < groups; ++gi) {
    OptTable t; int chunks;
    const int c = (count - done) < VQA_OPTIM_MAX_TENSORS ? (count - done) : VQA_OPTIM_MAX_TENSORS;
    if (int rc = build_table(ts + done, c, false, t, chunks)) return rc;
    const int grid = chunks < per_group ? (chunks > 0 ? chunks : 1) : per_group;
    grad_norm_sq_kernel<<<grid, kOptThreads, 0, s>>>(t, chunks, ws + used);
    VQA_LAUNCH_CHECK();
    used += grid; done += c;
  }
  clip_scale_kernel<<<1, 32, 0, s>>>(ws, used, max_norm, total_norm, scale);
  VQA_LAUNCH_CHECK();
  if (scale_in_place) {
    done = 0;
    for (int gi = 0; gi < groups; ++gi) {
      OptTable t; int chunks;
      const int c = (count - done) < VQA_OPTIM_MAX_TENSORS ? (count - done) : VQA_OPTIM_MAX_TENSORS;
      if (int rc = build_table(ts + done, c, false, t, chunks)) return rc;
      grad_scale_kernel<<<grid_for_chunks(chunks), kOptThreads, 0, s>>>(t, chunks, scale);
      VQA_LAUNCH_CHECK();
      done += c;
    }
  }
  return VQA_OK;
}

int adamax_step(const vqa_optim_tensor* ts, int count, float beta1, float beta2, float eps, float weight_decay, int step,
                const float* grad_scale, cudaStream_t s) {
  VQA_REQUIRE(step >= 1, "adamax_step: step %d must be >= 1", step);
  const double bias_correction = 1.0 - pow((double)beta1, (double)step);
  int done = 0;
  while (done < count) {
    OptTable t; int chunks;
    const int c = (count - done) < VQA_OPTIM_MAX_TENSORS ? (count - done) : VQA_OPTIM_MAX_TENSORS;
    if (int rc = build_table(ts + done, c, true, t, chunks)) return rc;
    for (int i = 0; i < c; ++i) t.lr[i] = (float)((double)ts[done + i].lr / bias_correction);
    adamax_kernel<<<grid_for_chunks(chunks), kOptThreads, 0, s>>>(t, chunks, beta1, beta2, eps, weight_decay, grad_scale);
    VQA_LAUNCH_CHECK();
    done += c;
  }
  return VQA_OK;
}

}  // namespace vqa
