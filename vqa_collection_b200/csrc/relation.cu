// k7 — spatial relation labels on device.
//
// Reference: util/relation.py:3-45 (spatial_relation) and :65-80
// (relation_graph).  Every unordered pair i<j is evaluated ONCE and yields both
// [i,j] (first return value) and [j,i] (second return value).
//
// Layout of the work: a persistent CTA walks over images; thread p of the CTA
// owns pair p = (i,j) for the whole kernel (K=36: 630 pairs -> 640 threads, the
// (i,j) decode happens once).  Per image the K boxes and their per-box terms
// (area, centre) go to shared memory once, each thread evaluates its pair with
// ~45 fp32 instructions and drops the two label bytes into a K×K shared tile
// that leaves with 16-byte coalesced stores.  Box tile and label tile are double
// buffered, so there is ONE __syncthreads per image and the box load of image
// n+1 is in flight under the pair evaluation of image n.
//
// Exactness (labels are bit-exact against the reference on the same float32 boxes):
//  * float32 geometry uses IEEE round-to-nearest intrinsics (no FMA contraction);
//  * `I == b` (relation.py:24) is four ordered compares: max(a,b)==b ⇔ a<=b;
//  * `iou >= 0.5` with iou = fl(ai/d) (relation.py:28-30) is decided without the
//    division: fl(ai/d) >= 0.5 ⇔ ai/d >= 0.5-2^-26 (ties-to-even lands on 0.5), i.e.
//    for d>0  ai - d/2 >= -d·2^-26.  ai - d/2 is exact when ai ∈ [d/4, d] (Sterbenz) and
//    otherwise so far from the threshold that its rounding cannot cross it; d<0 mirrors,
//    d==0 gives ±inf/NaN -> ai>0;
//  * `‖c_a-c_b‖/‖(w,h)‖ <= 0.5` (relation.py:37-38; float32 norm, float64 divide):
//    (double)nrm/D <= 0.5 ⇔ (double)nrm <= D/2 (exact: x > D/2 implies x/D >= 0.5+2^-53),
//    and since fl(sqrt(s)) is monotone in s the test is `s <= s_max` with s_max the
//    largest float32 whose correctly rounded root is <= D/2 (found once per image size);
//  * the direction octant is decided by exact sign/magnitude comparisons of the centre
//    offset instead of atan2f: away from and exactly on the 8 angular
//    boundaries this equals the reference's float32 arctan2 → rad2deg → -90 → %360 → /45
//    → ceil pipeline.
// Inputs are assumed finite (pixel coordinates).
#include <math.h>

#include "common.cuh"

namespace vqa {

constexpr int kRelMaxK = 64;
constexpr int kRelMaxThreads = 1024;
// pairs per thread: 1 up to K = 44 (946 pairs), 2 or 3 up to K = 64 (2016 pairs)

// largest float32 s with (double)fl(sqrt(s)) <= half_diag; -1 if there is none
__host__ __device__ __noinline__ float near_threshold_sq(double half_diag) {
  if (!(half_diag >= 0.0)) return -1.f;
  if (half_diag > 1.8e19) return 3.4028234663852886e38f;       // sqrt(FLT_MAX): everything finite is near
  float hf = (float)half_diag;                                 // round to nearest, then step down if above
  if ((double)hf > half_diag) hf = nextafterf(hf, -1.f);
  float s = hf * hf;
  for (int it = 0; it < 64 && (double)sqrtf(s) > (double)hf; ++it) s = nextafterf(s, -1.f);
  for (int it = 0; it < 64; ++it) {
    const float up = nextafterf(s, 3.4028234663852886e38f);
    if (up == s || (double)sqrtf(up) > (double)hf) break;
    s = up;
  }
  return s;
}

// Direction labels (ab | ba << 8) as a function of six predicates of the centre offset e = c_b - c_a:
//   bit 0: ex > 0   bit 1: ex < 0   bit 2: ey > 0   bit 3: ey < 0   bit 4: |ey| <= |ex|   bit 5: |ex| <= |ey|
// Branch-free octant: quadrant qn (counted from the first octant of label 4) and which half of it.
//   q0: ex>=0, ey<0  -> 4 | 5      q1: ex<0, ey<=0 -> 6 | 7      q2: ex<=0, ey>0 -> 8 | 9      q3: ex>0, ey>=0 -> 10 | 11
// second half of q0/q2 when !(|ey|<=|ex|), of q1/q3 when !(|ex|<=|ey|); the opposite direction is the quadrant
// two further on.  Three exceptions: θ=90 (ex>0, ey==0) is m=0 -> label 3, θ=-90 likewise for the way back,
// and coincident centres give atan2(0,0)=0 -> (9,5).
// The kernel evaluates this ONCE per combination into a 64-entry shared-memory table (kRelDirLut) and the pair loop
// looks the labels up: 6 compares + index arithmetic + one 16-bit load instead of ~35 instructions per pair.
__host__ __device__ __forceinline__ uint32_t relation_dir_labels(uint32_t bits) {
  const bool xp = bits & 1u, xn = bits & 2u, yp = bits & 4u, yn = bits & 8u, yx = bits & 16u, xy = bits & 32u;
  const uint32_t qn = (!xn & yn) ? 0u : (xn & !yp) ? 1u : (!xp & yp) ? 2u : 3u;
  const uint32_t sec = (qn & 1u) ? (uint32_t)!xy : (uint32_t)!yx;
  uint32_t lab_ab = 4u + 2u * qn + sec, lab_ba = 4u + 2u * ((qn + 2u) & 3u) + sec;
  const bool y0 = !yp & !yn;
  lab_ab = (y0 & xp) ? 3u : lab_ab;
  lab_ba = (y0 & xn) ? 3u : lab_ba;
  const bool both0 = y0 & !xp & !xn;
  lab_ab = both0 ? 9u : lab_ab;
  lab_ba = both0 ? 5u : lab_ba;
  return lab_ab | (lab_ba << 8);
}

// both labels of one unordered pair, packed as ab | ba << 8
__device__ __forceinline__ uint32_t relation_pair(const float4 a, const float4 b, const float4 ea,
                                                  const float4 eb, const float s_max, const uint16_t* __restrict__ dir_lut) {
  // e = (area, cx, cy, -)
  const bool inside = (a.x <= b.x) & (a.y <= b.y) & (b.z <= a.z) & (b.w <= a.w);    // I == b   (:24)
  const bool covered = (b.x <= a.x) & (b.y <= a.y) & (a.z <= b.z) & (a.w <= b.w);   // I == a   (:25)
  // IoU with UNCLAMPED areas (:28-30) — negative×negative is positive (F6)
  const float i0 = fmaxf(a.x, b.x), i1 = fmaxf(a.y, b.y);
  const float i2 = fminf(a.z, b.z), i3 = fminf(a.w, b.w);
  const float ai = __fmul_rn(__fsub_rn(i3, i1), __fsub_rn(i2, i0));
  const float d = __fsub_rn(__fadd_rn(ea.x, eb.x), ai);
  const float x = __fsub_rn(ai, __fmul_rn(0.5f, d));
  const float t = __fmul_rn(d, -0x1p-26f);
  const bool overlap = ((d > 0.f) & (x >= t)) | ((d < 0.f) & (x <= t)) | ((d == 0.f) & (ai > 0.f));
  // centre distance (:33-38)
  const float dx = __fsub_rn(ea.y, eb.y), dy = __fsub_rn(ea.z, eb.z);
  const float s = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
  const bool near = s <= s_max;
  // direction of b seen from a (:39-42): θ = atan2(ex, ey), δ = θ-90, label = ceil((δ mod 360)/45)+3
  // and the same for δ+180.  e = c_b - c_a = -(c_a - c_b) exactly; the labels come out of the table (see above)
  const float ex = -dx, ey = -dy;
  const uint32_t bits = (uint32_t)(ex > 0.f) | ((uint32_t)(ex < 0.f) << 1) | ((uint32_t)(ey > 0.f) << 2) |
                        ((uint32_t)(ey < 0.f) << 3) | ((uint32_t)(fabsf(ey) <= fabsf(ex)) << 4) |
                        ((uint32_t)(fabsf(ex) <= fabsf(ey)) << 5);
  const uint32_t dir = dir_lut[bits];
  return inside ? (1u | (2u << 8)) : covered ? (2u | (1u << 8)) : overlap ? (3u | (3u << 8)) : near ? dir : 0u;
}

// Thread roles: the first `pair_threads` threads own one (or two) box pairs each and do nothing but evaluate them; the
// last kRelIoThreads threads (one warp) load and stage the next image's boxes and write the finished label tile out.
// (With one role for all, every pair thread issued the predicated-off box staging and the tile-store loop as well:
// 168 instructions per image and thread, of which ~100 were the pair itself; profiles/r01e_ncu_relation.md.)
constexpr int kRelIoThreads = 32;

// kMaxT / kMinCtas: launch bounds of the instantiation (K = 36: 640 pair threads + 32, three CTAs per SM)
template <int kRelPairsPerThread, int kMaxT, int kMinCtas>
__global__ void __launch_bounds__(kMaxT, kMinCtas)
relation_labels_kernel(const float4* __restrict__ bbox, const float2* __restrict__ wh, int B, int K,
                       float s_max_uniform, uint8_t* __restrict__ labels, int pair_threads) {
  __shared__ float4 s_box[2][kRelMaxK];
  __shared__ float4 s_ext[2][kRelMaxK];
  __shared__ float s_thr[2];
  __shared__ uint16_t s_dir[64];
  __shared__ __align__(16) uint8_t s_lab[2][kRelMaxK * kRelMaxK];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int KK = K * K, npairs = K * (K - 1) / 2;
  const bool io = tid >= pair_threads;                // warp-uniform: pair_threads is a multiple of 32
  const int io_tid = tid - pair_threads;
  // this thread's pairs, decoded once: p -> (i,j), i<j, row-major over the upper triangle
  int pi[kRelPairsPerThread], pj[kRelPairsPerThread];
#pragma unroll
  for (int u = 0; u < kRelPairsPerThread; ++u) {
    int p = tid + u * pair_threads, i = 0;
    pi[u] = -1; pj[u] = 0;
    if (!io && p < npairs) {
      while (p >= K - 1 - i) { p -= K - 1 - i; ++i; }
      pi[u] = i; pj[u] = i + 1 + p;
    }
  }
  for (int t = tid; t < 2 * kRelMaxK * kRelMaxK; t += nthr) (&s_lab[0][0])[t] = 0;    // the diagonal stays 0
  if (tid < 64) s_dir[tid] = (uint16_t)relation_dir_labels((uint32_t)tid);

  auto stage_boxes = [&](int img, int buf, int k, float4 bx) {
    // per-box terms (:28-29,:33-35): area = (y1-y0)(x1-x0), centre = x0 + (x1-x0)/2
    const float w = __fsub_rn(bx.z, bx.x), h = __fsub_rn(bx.w, bx.y);
    s_box[buf][k] = bx;
    s_ext[buf][k] = make_float4(__fmul_rn(h, w), __fadd_rn(bx.x, __fmul_rn(w, 0.5f)),
                                     __fadd_rn(bx.y, __fmul_rn(h, 0.5f)), 0.f);
    if (k == 0 && wh != nullptr) {
      const float2 dd = __ldg(wh + img);
      s_thr[buf] = near_threshold_sq(0.5 * sqrt((double)dd.x * (double)dd.x + (double)dd.y * (double)dd.y));
    }
  };

  int img = blockIdx.x, buf = 0;
  if (io && img < B)
    for (int k = io_tid; k < K; k += kRelIoThreads) stage_boxes(img, 0, k, __ldg(bbox + (size_t)img * K + k));
  __syncthreads();
  if (!io) {
    // ===== pair threads =====
    const int st_threads = pair_threads < 96 ? pair_threads : 96;      // the warps that also write the tile out
    for (; img < B; img += gridDim.x, buf ^= 1) {
      const float s_max = wh != nullptr ? s_thr[buf] : s_max_uniform;
#pragma unroll
      for (int u = 0; u < kRelPairsPerThread; ++u) {
        if (pi[u] >= 0) {
          const int i = pi[u], j = pj[u];
          const uint32_t l = relation_pair(s_box[buf][i], s_box[buf][j], s_ext[buf][i], s_ext[buf][j], s_max, s_dir);
          s_lab[buf][i * K + j] = (uint8_t)(l & 0xffu);
          s_lab[buf][j * K + i] = (uint8_t)(l >> 8);
        }
      }
      __syncthreads();
      // the finished tile leaves through the first pair warps (a warp-uniform branch; with the single I/O warp doing it,
      // that warp's LDS -> STG round trips were the longest path of an image) while the CTA already works on image n+1:
      // the tile is rewritten two images later, after the next barrier
      if (tid < st_threads) {
        uint8_t* out = labels + (size_t)img * KK;
        if ((KK & 15) == 0) {           // 36*36 = 1296 = 81 * 16: vector stores
          const uint4* src = reinterpret_cast<const uint4*>(s_lab[buf]);
          uint4* dst = reinterpret_cast<uint4*>(out);
          for (int t = tid; t < KK / 16; t += st_threads) __stcs(dst + t, src[t]);
        } else {
          for (int t = tid; t < KK; t += st_threads) out[t] = s_lab[buf][t];
        }
      }
    }
  } else {
    // ===== I/O threads: boxes in (loaded TWO images ahead: a DRAM round trip is longer than one image's pair math at
    // three CTAs per SM), tile of image n out =====
    static_assert(kRelMaxK <= 2 * kRelIoThreads, "two boxes per I/O thread");
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_boxes = [&](int im, float4& b0, float4& b1) {
      b0 = (im < B && io_tid < K) ? __ldg(bbox + (size_t)im * K + io_tid) : zero4;
      b1 = (im < B && io_tid + kRelIoThreads < K) ? __ldg(bbox + (size_t)im * K + io_tid + kRelIoThreads) : zero4;
    };
    float4 n0, n1;                                     // boxes of image n+1, in registers
    load_boxes(img + gridDim.x, n0, n1);
    for (; img < B; img += gridDim.x, buf ^= 1) {
      const int nxt = img + gridDim.x;
      float4 f0, f1;                                   // boxes of image n+2: in flight across this iteration
      load_boxes(nxt + gridDim.x, f0, f1);
      if (nxt < B) {
        if (io_tid < K) stage_boxes(nxt, buf ^ 1, io_tid, n0);
        if (io_tid + kRelIoThreads < K) stage_boxes(nxt, buf ^ 1, io_tid + kRelIoThreads, n1);
      }
      n0 = f0; n1 = f1;
      __syncthreads();
    }
  }
}

float relation_near_threshold(float img_w, float img_h) {
  return near_threshold_sq(0.5 * sqrt((double)img_w * (double)img_w + (double)img_h * (double)img_h));
}

int relation_labels(const float* d_bbox, const float* d_wh, int B, int K, float img_w, float img_h,
                    uint8_t* d_labels, cudaStream_t s) {
  VQA_REQUIRE(K >= 1 && K <= kRelMaxK, "relation_labels: K=%d out of range [1,%d]", K, kRelMaxK);
  VQA_REQUIRE(B >= 0, "relation_labels: B=%d", B);
  if (B == 0) return VQA_OK;                       // empty batch: nothing to do (pointers may be NULL)
  VQA_REQUIRE(d_bbox && d_labels, "relation_labels: NULL pointer");
  const double half_diag = 0.5 * sqrt((double)img_w * (double)img_w + (double)img_h * (double)img_h);
  const int npairs = K * (K - 1) / 2;
  // pairs per thread: the fewest with which the pair threads + the two I/O warps fit a block
  int ppt = 1;
  while (((npairs + ppt - 1) / ppt + 31) / 32 * 32 + kRelIoThreads > kRelMaxThreads) ++ppt;
  int pair_threads = ((npairs + ppt - 1) / ppt + 31) / 32 * 32;
  if (pair_threads < 32) pair_threads = 32;           // K = 1: no pairs, but the first pair warp still writes the tile out
  const int threads = pair_threads + kRelIoThreads;  // + one warp that moves boxes in and label tiles out
  VQA_REQUIRE(ppt <= 3, "relation_labels: K=%d needs %d pairs per thread", K, ppt);
  auto kernel = (ppt == 1 && threads <= 672) ? relation_labels_kernel<1, 672, 3>
              : ppt == 1 ? relation_labels_kernel<1, kRelMaxThreads, 1>
              : ppt == 2 ? relation_labels_kernel<2, kRelMaxThreads, 1> : relation_labels_kernel<3, kRelMaxThreads, 1>;
  static int ctas_per_sm[kRelMaxThreads / 32 + 1] = {0};
  int& occ = ctas_per_sm[threads / 32];
  if (occ == 0) {
    VQA_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, 0));
    if (occ < 1) occ = 1;
  }
  const long long resident = (long long)sm_count() * occ;
  const int grid = (int)(B < resident ? B : resident);
  kernel<<<grid, threads, 0, s>>>(reinterpret_cast<const float4*>(d_bbox), reinterpret_cast<const float2*>(d_wh), B, K,
                                  near_threshold_sq(half_diag), d_labels, pair_threads);
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
