// k7 — spatial relation labels on device.
//
// Reference: util/relation.py:3-45 (spatial_relation) and :65-80
// (relation_graph).  One CTA per image; every unordered pair i<j is evaluated
// ONCE and writes both [i,j] (first return value) and [j,i] (second return
// value) into a shared-memory K×K byte tile that is then stored with coalesced
// 16-byte writes.  float32 geometry is mirrored with IEEE round-to-nearest
// intrinsics (no FMA contraction); the direction octant is decided by exact
// sign/magnitude comparisons of the centre offset instead of atan2f (SURVEY.md
// H2): away from and exactly on the 8 angular boundaries this equals the
// reference's float32 arctan2 → rad2deg → -90 → %360 → /45 → ceil pipeline.
#include "common.cuh"

namespace vqa {

struct LabelPair { uint8_t ab, ba; };

__device__ __forceinline__ LabelPair relation_pair(const float4 a, const float4 b,
                                                   const double half_diag) {
  // intersection box (relation.py:19-22); float4 = (x0,y0,x1,y1)
  const float i0 = fmaxf(a.x, b.x), i1 = fmaxf(a.y, b.y);
  const float i2 = fminf(a.z, b.z), i3 = fminf(a.w, b.w);
  if (i0 == b.x && i1 == b.y && i2 == b.z && i3 == b.w) return {1, 2};   // :24
  if (i0 == a.x && i1 == a.y && i2 == a.z && i3 == a.w) return {2, 1};   // :25
  // IoU with UNCLAMPED areas (:28-30) — negative×negative is positive (F6)
  const float ai = __fmul_rn(__fsub_rn(i3, i1), __fsub_rn(i2, i0));
  const float aa = __fmul_rn(__fsub_rn(a.w, a.y), __fsub_rn(a.z, a.x));
  const float ab = __fmul_rn(__fsub_rn(b.w, b.y), __fsub_rn(b.z, b.x));
  const float iou = __fdiv_rn(ai, __fsub_rn(__fadd_rn(aa, ab), ai));
  if (iou >= 0.5f) return {3, 3};
  // centres (:33-35): x0 + (x1-x0)/2
  const float cax = __fadd_rn(a.x, __fmul_rn(__fsub_rn(a.z, a.x), 0.5f));
  const float cay = __fadd_rn(a.y, __fmul_rn(__fsub_rn(a.w, a.y), 0.5f));
  const float cbx = __fadd_rn(b.x, __fmul_rn(__fsub_rn(b.z, b.x), 0.5f));
  const float cby = __fadd_rn(b.y, __fmul_rn(__fsub_rn(b.w, b.y), 0.5f));
  const float dx = __fsub_rn(cax, cbx), dy = __fsub_rn(cay, cby);
  const float nrm = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
  // (double)nrm / ‖(w,h)‖ <= 0.5  ⇔  (double)nrm <= ‖(w,h)‖/2  (exact)
  if (!((double)nrm <= half_diag)) return {0, 0};                        // :37-38,45
  // direction of b seen from a (:39-42): θ = atan2(ex, ey), δ = θ-90,
  // label = ceil((δ mod 360)/45)+3 and the same for δ+180.
  const float ex = __fsub_rn(cbx, cax), ey = __fsub_rn(cby, cay);
  if (ex > 0.f) {
    if (ey == 0.f) return {3, 7};                       // θ = 90  → m = 0  (label-3 quirk)
    if (ey < 0.f) return (-ey <= ex) ? LabelPair{4, 8} : LabelPair{5, 9};   // (90,135] | (135,180)
    return (ex <= ey) ? LabelPair{10, 6} : LabelPair{11, 7};                // (0,45] | (45,90)
  }
  if (ex < 0.f) {
    if (ey == 0.f) return {7, 3};                       // θ = -90 → δ+180 = 0 → 3
    if (ey < 0.f) return (ex >= ey) ? LabelPair{6, 10} : LabelPair{7, 11};  // (-180,-135] | (-135,-90)
    return (ey <= -ex) ? LabelPair{8, 4} : LabelPair{9, 5};                 // (-90,-45] | (-45,0)
  }
  // ex == 0: straight up/down, or coincident centres (atan2(0,0) = 0 → 9,5)
  return (ey < 0.f) ? LabelPair{5, 9} : LabelPair{9, 5};
}

constexpr int kRelThreads = 256;
constexpr int kRelMaxK = 64;

__global__ void __launch_bounds__(kRelThreads)
relation_labels_kernel(const float4* __restrict__ bbox, const float2* __restrict__ wh,
                       int B, int K, double half_diag_uniform, uint8_t* __restrict__ labels) {
  __shared__ float4 s_box[kRelMaxK];
  __shared__ __align__(16) uint8_t s_lab[kRelMaxK * kRelMaxK];
  const int KK = K * K;
  for (int img = blockIdx.x; img < B; img += gridDim.x) {
    if (threadIdx.x < K) s_box[threadIdx.x] = __ldg(bbox + (size_t)img * K + threadIdx.x);
    double half_diag = half_diag_uniform;
    if (wh != nullptr) {
      const float2 d = __ldg(wh + img);
      half_diag = 0.5 * sqrt((double)d.x * (double)d.x + (double)d.y * (double)d.y);
    }
    __syncthreads();
    for (int t = threadIdx.x; t < KK; t += kRelThreads) {
      const int i = t / K, j = t - i * K;
      if (i == j) s_lab[t] = 0;
      if (i < j) {
        const LabelPair l = relation_pair(s_box[i], s_box[j], half_diag);
        s_lab[t] = l.ab;
        s_lab[j * K + i] = l.ba;
      }
    }
    __syncthreads();
    uint8_t* out = labels + (size_t)img * KK;
    if ((KK & 15) == 0) {           // 36*36 = 1296 = 81 * 16: vector stores
      const uint4* src = reinterpret_cast<const uint4*>(s_lab);
      uint4* dst = reinterpret_cast<uint4*>(out);
      for (int t = threadIdx.x; t < KK / 16; t += kRelThreads) dst[t] = src[t];
    } else {
      for (int t = threadIdx.x; t < KK; t += kRelThreads) out[t] = s_lab[t];
    }
    __syncthreads();
  }
}

int relation_labels(const float* d_bbox, const float* d_wh, int B, int K, float img_w, float img_h,
                    uint8_t* d_labels, cudaStream_t s) {
  VQA_REQUIRE(K >= 1 && K <= kRelMaxK, "relation_labels: K=%d out of range [1,%d]", K, kRelMaxK);
  VQA_REQUIRE(B >= 0, "relation_labels: B=%d", B);
  if (B == 0) return VQA_OK;                       // empty batch: nothing to do (pointers may be NULL)
  VQA_REQUIRE(d_bbox && d_labels, "relation_labels: NULL pointer");
  const double half_diag = 0.5 * sqrt((double)img_w * (double)img_w + (double)img_h * (double)img_h);
  const int grid = B < sm_count() * 8 ? B : sm_count() * 8;
  relation_labels_kernel<<<grid, kRelThreads, 0, s>>>(reinterpret_cast<const float4*>(d_bbox),
                                                      reinterpret_cast<const float2*>(d_wh), B, K,
                                                      half_diag, d_labels);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
