// Persistent fused GRU for the question encoder (bf16 operands, sm_100a).
//
// Reference: modules/modules.py:139-159 (SentenceEmbedding: nn.GRU, 1 layer, h0 = 0,
// output[:, -1]) fed by encoder.py:159 (embedding).  torch GRU semantics:
//   r = σ(W_ir x + b_ir + W_hr h + b_hr)      z = σ(W_iz x + b_iz + W_hz h + b_hz)
//   n = tanh(W_in x + b_in + r ⊙ (W_hn h + b_hn))      h' = (1 - z) ⊙ n + z ⊙ h
//
// ONE cooperative launch runs all T steps.  CTA (m, j) owns 128 batch rows × 64 hidden
// units for the whole sequence:
//   * weights are pre-packed so that the j-th 192-row block of Wx/Wh holds [r | z | n]
//     rows of units [64j, 64j+64): the three gates of a unit land in the same CTA
//   * TMEM accumulators (256 fp32 columns, double buffered over time steps):
//       [0,64) r   [64,128) z   [128,192) W_in x + W_hn h   [192,256) W_in x
//     the x-part of step t+1 (independent of h) is issued while step t's epilogue and
//     the grid barrier are still in flight.  The h-part is ONE N=192 MMA per k-step on top of
//     the x-part's [r|z|n] tile (the A tile is read from shared memory once — the kernel is
//     shared-memory-bandwidth bound: TMA fills + MMA operand reads ≈ 2 MB per step per SM);
//     the epilogue recovers W_hn h as column [128,192) minus the separately kept W_in x
//   * epilogue threads (one accumulator row, 16 units each) keep their fp32 state values in
//     REGISTERS across all steps; only the bf16 copy that feeds the next step's MMA goes
//     to global memory (double buffered), followed by a grid-wide arrive/wait on a
//     global counter (release/acquire + async-proxy fence, since TMA reads it)
//   * no gi/gh round trip, no per-step launches: 1 launch instead of 2T+1.
//   * the CTAs of one row block all read the same x_t / h_{t-1} tiles; h is freshly written every
//     step (its L2 lines cannot be served from a read-only replica); measured, those reads cost
//     more than the larger W tiles.  Optionally (VQA_B200_GRU_CLUSTER=2|4|8) the CTAs form
//     thread-block clusters along the unit dimension: each CTA loads a 1/cs slice of the x / h
//     tile and TMA-multicasts it to its cluster, and a stage is recycled only when every CTA of
//     the cluster has consumed it (tcgen05.commit multicast onto all the empty barriers).
#include <stdlib.h>

#include "tc_common.cuh"

namespace vqa {
namespace gru {

using namespace tc;

constexpr int UNITS = 64;
constexpr int WROWS = 3 * UNITS;                  // 192
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 16;                     // 4 warps per TMEM lane quarter, 16 units each: the gate math is
                                                  // latency-bound (dependent MUFU chains), more warps hide it
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;   // 640
constexpr int UPT = UNITS / (EPI_WARPS / 4);      // units per epilogue thread = 16
constexpr int W_PREFETCH = 4;                     // W_h k-blocks issued ahead of the grid barrier
constexpr int A_BYTES = BM * BK * 2;              // 16 KB
constexpr int W_BYTES = WROWS * BK * 2;           // 24 KB
constexpr int STAGE_BYTES = A_BYTES + W_BYTES;    // 40 KB
constexpr int STAGES = 5;
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;
constexpr int COL_R = 0, COL_Z = 64, COL_NH = 128, COL_NI = 192;
constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256 + 4 * UNITS * 4;

struct Params {
  int B, T, H, E_pad, tiles_n, num_ctas;
  int cs;                       // cluster size along the unit dimension (1, 2, 4 or 8)
  const float* bias;            // packed [4H]: b_ir+b_hr | b_iz+b_hz | b_in | b_hn
  __nv_bfloat16* h_op[2];       // bf16 state, double buffered over steps
  float* h_last;                // [B,H] f32 or NULL
  __nv_bfloat16* h_last_lp;     // [B,H] bf16 or NULL
  __nv_bfloat16* h_all;         // [B,T,H] bf16 or NULL: every state; then it also IS the operand buffer
                                // (both tensor maps view it as [B, T*H], step t reads columns (t-1)*H..)
  int* counter;                 // per-row-block arrival counters [tiles_m], zero on entry
  int debug;                    // timing experiments only (VQA_B200_GRU_DEBUG): 1 no barrier wait, 2 no gate math, 4 no fence, 8 no h-tile loads, 16 no W_h-tile loads
};

// Gate non-linearities on tanh.approx.f32: 1 MUFU each (ex2+rcp forms cost 2 and made the epilogue MUFU-bound:
// measured 212 -> 202 us at B=1024, T=14); max relative error 2^-11, below the bf16 rounding (2^-9) that the
// state operand of the next step's MMA gets anyway.  This kernel only serves the bf16 mode; fp32 mode runs the
// per-step path with tanhf/expf (pool.cu).
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

__global__ void __launch_bounds__(THREADS, 1)
gru_persistent_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH0,
                      const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmWx,
                      const __grid_constant__ CUtensorMap tmWh, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 4));
  float* bias_s = reinterpret_cast<float*>(base_ptr + STAGES * STAGE_BYTES + 256);   // [4][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_blk = blockIdx.x / p.tiles_n, n_blk = blockIdx.x % p.tiles_n;
  const int m0 = m_blk * BM, u0 = n_blk * UNITS;
  const int kb_x = p.E_pad / BK, kb_h = p.H / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmH0); tma_prefetch_desc(&tmH1);
    tma_prefetch_desc(&tmWx); tma_prefetch_desc(&tmWh);
  }
  const int cs = p.cs;
  const uint32_t crank = cs > 1 ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << cs) - 1u);
  const int slice_rows = BM / cs;                      // rows of the x / h tile this CTA loads for its cluster
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), cs); }
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), EPI_THREADS); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp == 3) {
    for (int i = lane; i < 4 * UNITS; i += 32) bias_s[i] = p.bias[(i / UNITS) * p.H + u0 + (i % UNITS)];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();                      // peers multicast into / arrive on these barriers
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();              // PDL primary only (cooperative launches are not made dependents)
  // A-operand tile: this CTA's row slice, delivered to every CTA of the cluster
  auto load_a = [&](uint32_t sa, const CUtensorMap* map, uint32_t bar, int col) {
    if (cs > 1) tma_load_2d_multicast(sa + crank * (slice_rows * 128), map, bar, col, m0 + (int)crank * slice_rows, cmask);
    else tma_load_2d(sa, map, bar, col, m0);
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int kb = 0; kb < kb_x; ++kb) {                       // x-part: no dependence on h
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), STAGE_BYTES);
          load_a(sa, &tmX, full_bar(stage), t * p.E_pad + kb * BK);
          tma_load_2d(sw, &tmWx, full_bar(stage), kb * BK, n_blk * WROWS);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (t > 0) {
          // W_h tiles do not depend on h: arm the first stages and start their W loads BEFORE the barrier
          const int npre = kb_h < W_PREFETCH ? kb_h : W_PREFETCH;
          int pre_stage[W_PREFETCH];
          for (int kb = 0; kb < npre; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            mbar_arrive_expect_tx(full_bar(stage), ((p.debug & 8) ? 0 : A_BYTES) + ((p.debug & 16) ? 0 : W_BYTES));
            if (!(p.debug & 16)) tma_load_2d(base + stage * STAGE_BYTES + A_BYTES, &tmWh, full_bar(stage), kb * BK, n_blk * WROWS);
            pre_stage[kb] = stage;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          // barrier over the tiles_n CTAs that own the same batch rows: h_{t-1}[rows, :] is published
          const int target = t * p.tiles_n;
          if (!(p.debug & 1)) while (ld_acquire_gpu(p.counter + m_blk) < target) { }
          fence_proxy_async_all();
          const CUtensorMap* tmH = ((t - 1) & 1) ? &tmH1 : &tmH0;
          const int hcol = p.h_all ? (t - 1) * p.H : 0;
          for (int kb = 0; kb < npre; ++kb)
            if (!(p.debug & 8)) load_a(base + pre_stage[kb] * STAGE_BYTES, tmH, full_bar(pre_stage[kb]), hcol + kb * BK);
          for (int kb = npre; kb < kb_h; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
            mbar_arrive_expect_tx(full_bar(stage), ((p.debug & 8) ? 0 : A_BYTES) + ((p.debug & 16) ? 0 : W_BYTES));
            if (!(p.debug & 8)) load_a(sa, tmH, full_bar(stage), hcol + kb * BK);
            if (!(p.debug & 16)) tma_load_2d(sw, &tmWh, full_bar(stage), kb * BK, n_blk * WROWS);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_rzn = make_idesc_bf16(BM, 3 * UNITS);    // N = 192: r,z,n rows
      constexpr uint32_t idesc_n = make_idesc_bf16(BM, UNITS);          // N = 64 : n rows
      constexpr uint32_t N_ROW_OFF = (2 * UNITS * BK * 2) >> 4;         // rows 128.. of the W tile (16 KB)
      int stage = 0; uint32_t phase = 0;
      for (int t = 0; t < p.T; ++t) {
        const int acc = t & 1;
        const uint32_t acc_phase = (uint32_t)(t >> 1) & 1u;
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < kb_x; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
          const uint64_t adesc = make_sw128_kmajor_desc(sa), wdesc = make_sw128_kmajor_desc(sw);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            umma_bf16(d + COL_R, adesc + 2 * k, wdesc + 2 * k, idesc_rzn, (kb | k) != 0);       // [r_x | z_x | n_x]
            umma_bf16(d + COL_NI, adesc + 2 * k, wdesc + N_ROW_OFF + 2 * k, idesc_n, (kb | k) != 0);   // n_x again, kept apart
          }
          if (cs > 1) umma_commit_multicast(empty_bar(stage), cmask); else umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (t > 0) {
          for (int kb = 0; kb < kb_h; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
            const uint64_t adesc = make_sw128_kmajor_desc(sa), wdesc = make_sw128_kmajor_desc(sw);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              umma_bf16(d + COL_R, adesc + 2 * k, wdesc + 2 * k, idesc_rzn, 1u);                  // one MMA, A tile read once
            }
            if (cs > 1) umma_commit_multicast(empty_bar(stage), cmask); else umma_commit(empty_bar(stage));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit(tfull_bar(acc));
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== gate epilogue: thread = (batch row, 32 units), fp32 state in registers =====
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int uh = (warp - EPI_WARP0) >> 2;          // which UPT-wide slice of the 64 units
    const int et = threadIdx.x - EPI_WARP0 * 32;
    const int row = m0 + q * 32 + lane;
    const bool row_ok = row < p.B;
    const int ub = uh * UPT;                         // first unit (within the tile) of this thread
    float h[UPT];
#pragma unroll
    for (int j = 0; j < UPT; ++j) h[j] = 0.f;
    for (int t = 0; t < p.T; ++t) {
      const int acc = t & 1;
      const uint32_t acc_phase = (uint32_t)(t >> 1) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16) + ub;
      const bool last = (t == p.T - 1);
      __nv_bfloat16* hdst = p.h_all ? p.h_all + (size_t)t * p.H : ((last && p.h_last_lp) ? p.h_last_lp : p.h_op[t & 1]);
      const size_t h_ld = p.h_all ? (size_t)p.T * p.H : (size_t)p.H;
#pragma unroll
      for (int c = 0; c < UPT; c += 8) {
        uint32_t vr[8], vz[8], vni[8], vnh[8];
        tmem_ld_32x8(trow + COL_R + c, vr);
        tmem_ld_32x8(trow + COL_Z + c, vz);
        tmem_ld_32x8(trow + COL_NI + c, vni);
        tmem_ld_32x8(trow + COL_NH + c, vnh);
        tmem_ld_wait();
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (p.debug & 2) { o[j] = __uint_as_float(vr[j]) + __uint_as_float(vz[j]) + __uint_as_float(vni[j]); h[c + j] = o[j]; continue; }
          const float pr = __uint_as_float(vr[j]) + bias_s[ub + c + j], pz = __uint_as_float(vz[j]) + bias_s[UNITS + ub + c + j];
          const float r = sigmoid_fast(pr), z = sigmoid_fast(pz);
          // columns COL_NH hold n_x + W_hn·h (x-part and h-part accumulate into one N=192 tile); n_x alone is in COL_NI
          const float nh = (__uint_as_float(vnh[j]) - __uint_as_float(vni[j])) + bias_s[3 * UNITS + ub + c + j];
          const float pn = __uint_as_float(vni[j]) + bias_s[2 * UNITS + ub + c + j] + r * nh;
          const float n = tanh_fast(pn);
          const float hn = (1.f - z) * n + z * h[c + j];
          h[c + j] = hn;
          o[j] = hn;
        }
        if (row_ok) {
          uint4 w0;
          w0.x = pack_bf16x2(o[0], o[1]); w0.y = pack_bf16x2(o[2], o[3]);
          w0.z = pack_bf16x2(o[4], o[5]); w0.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(hdst + (size_t)row * h_ld + u0 + ub + c) = w0;
        }
      }
      tcgen05_fence_before();
      mbar_arrive(tempty_bar(acc));
      if (!last) {
        // publish h_t: make the stores visible GPU-wide, then one arrive per CTA on its row-block counter
        // (a per-warp publish — 8 arrivals per CTA, no bar.sync — measured the same: 229 vs 226 us)
        if (!(p.debug & 4)) __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (et == 0) {
          fence_proxy_async_all();
          red_release_gpu_add(p.counter + m_blk, 1);
        }
      }
    }
    if (row_ok && p.h_last) {
      float* dst = p.h_last + (size_t)row * p.H + u0 + ub;
#pragma unroll
      for (int j = 0; j < UPT; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(h[j], h[j + 1], h[j + 2], h[j + 3]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (cs > 1) cluster_sync_all();                      // no CTA leaves while a peer may still signal its barriers
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace gru

// X [B*T, E_pad] bf16 (row b*T+t); wx_p [3H,E_pad], wh_p [3H,H] packed; bias_p [4H];
// h_op: 2*B*H bf16 scratch; counter: >= 64 ints (zeroed here).
int gru_persistent(const void* X, int B, int T, int H, int E_pad, const void* wx_p, const void* wh_p,
                   const float* bias_p, void* h_op, int* counter, float* h_last, void* h_last_lp, void* h_all,
                   cudaStream_t s) {
  using namespace gru;
  VQA_REQUIRE(H % UNITS == 0 && E_pad % tc::BK == 0, "gru(bf16): H=%d must be a multiple of 64 and E_pad=%d of 64", H, E_pad);
  const int tiles_n = H / UNITS;
  const int sms = sm_count();
  VQA_REQUIRE(tiles_n <= sms, "gru(bf16): H=%d needs more CTAs than the device has SMs", H);
  const int max_tiles_m = sms / tiles_n;
  static DeviceOnce attr;                            // per device, not per process
  if (attr.need(current_device())) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(gru_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr.mark(current_device());
  }
  CUtensorMap tmWx, tmWh;
  int rc;
  static int max_cs = 8;                               // halved whenever a cluster launch of that size is refused
  int cs = 1;
  // Measured at B=1024: with the grid barrier disabled multicast saves 21 us of operand traffic
  // (172 -> 151 us), but the real step is bound by the publish -> acquire -> load latency chain and comes out the
  // same (212 us at cluster sizes 1, 2, 4 and 8).  Default = plain launch; VQA_B200_GRU_CLUSTER=N turns it on.
  int want = 1;
  if (const char* e = getenv("VQA_B200_GRU_CLUSTER")) { const int v = atoi(e); if (v >= 1) want = v < max_cs ? v : max_cs; }
  while (cs < want && tiles_n % (cs * 2) == 0) cs *= 2;
  if ((rc = tc::make_tensor_map_bf16(&tmWx, wx_p, 3LL * H, E_pad, E_pad, WROWS))) return rc;
  if ((rc = tc::make_tensor_map_bf16(&tmWh, wh_p, 3LL * H, H, H, WROWS))) return rc;
  // batch chunks so that every chunk's CTAs are co-resident (grid barrier)
  for (int b0 = 0; b0 < B; b0 += max_tiles_m * tc::BM) {
    const int Bc = (B - b0 < max_tiles_m * tc::BM) ? B - b0 : max_tiles_m * tc::BM;
    const int tiles_m = (Bc + tc::BM - 1) / tc::BM;
    const __nv_bfloat16* Xc = (const __nv_bfloat16*)X + (size_t)b0 * T * E_pad;
    __nv_bfloat16* h0 = (__nv_bfloat16*)h_op + (size_t)b0 * H;
    __nv_bfloat16* h1 = (__nv_bfloat16*)h_op + (size_t)B * H + (size_t)b0 * H;
    CUtensorMap tmX, tmH0, tmH1;
    if ((rc = tc::make_tensor_map_bf16(&tmX, Xc, Bc, (long long)T * E_pad, (long long)T * E_pad, tc::BM / cs))) return rc;
    __nv_bfloat16* hall = h_all ? (__nv_bfloat16*)h_all + (size_t)b0 * T * H : nullptr;
    if (hall) {
      if ((rc = tc::make_tensor_map_bf16(&tmH0, hall, Bc, (long long)T * H, (long long)T * H, tc::BM / cs))) return rc;
      tmH1 = tmH0;
    } else {
      if ((rc = tc::make_tensor_map_bf16(&tmH0, h0, Bc, H, H, tc::BM / cs))) return rc;
      if ((rc = tc::make_tensor_map_bf16(&tmH1, h1, Bc, H, H, tc::BM / cs))) return rc;
    }
    Params p;
    p.B = Bc; p.T = T; p.H = H; p.E_pad = E_pad; p.tiles_n = tiles_n; p.num_ctas = tiles_m * tiles_n;
    p.bias = bias_p; p.h_op[0] = h0; p.h_op[1] = h1;
    p.h_last = h_last ? h_last + (size_t)b0 * H : nullptr;
    p.h_all = hall;
    p.h_last_lp = h_last_lp ? (__nv_bfloat16*)h_last_lp + (size_t)b0 * H : nullptr;
    p.counter = counter;
    p.cs = cs;
    { const char* e = getenv("VQA_B200_GRU_DEBUG"); p.debug = e ? atoi(e) : 0; }
    VQA_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int) * tiles_m, s));
    void* args[] = {&tmX, &tmH0, &tmH1, &tmWx, &tmWh, &p};
    if (cs > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(p.num_ctas); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
      cudaLaunchAttribute at[2];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
      cfg.attrs = at; cfg.numAttrs = 2;
      cudaError_t e = cudaLaunchKernelExC(&cfg, (const void*)gru_persistent_kernel, args);
      if (e != cudaSuccess) {
        // clusters of this size cannot all be co-resident on this device: remember, retry the whole call smaller
        (void)cudaGetLastError();
        max_cs = cs / 2;
        return gru_persistent(X, B, T, H, E_pad, wx_p, wh_p, bias_p, h_op, counter, h_last, h_last_lp, h_all, s);
      }
    } else {
      VQA_CUDA_CHECK(cudaLaunchCooperativeKernel((const void*)gru_persistent_kernel, dim3(p.num_ctas), dim3(THREADS), args,
                                                 (size_t)SMEM_BYTES, s));
    }
    count_launch();
  }
  return VQA_OK;
}

}  // namespace vqa
