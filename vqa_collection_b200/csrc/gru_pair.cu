// Persistent fused GRU on CTA PAIRS (tcgen05 cta_group::2) — same algorithm, outputs and packed
// weights as gru_tc.cu (read that header first); what changes is who holds which operand.
//
// gru_tc.cu is bound by shared-memory bandwidth per SM: every step each CTA fills 855 KB through
// TMA and the MMAs read 0.96 MB of operands back out.  Here two CTAs on one TPC form a pair that
// owns 256 batch rows × 64 hidden units:
//   * each CTA still owns (loads, gates, publishes) its own 128 rows — the epilogue, the fp32 state
//     in registers and the per-row-block publish/acquire protocol are unchanged;
//   * the 192-row W tile of a k-block is SPLIT: CTA r loads rows [r|z|n] of units [32r, 32r+32)
//     (three 32-row TMA boxes out of the same gate-interleaved packing) — half the W fill per SM;
//   * ONE thread of the pair's leader CTA issues tcgen05.mma.cta_group::2 (M = 256, N = 192): each
//     SM's tensor core reads its own A tile and both B halves, so the B operand reads per SM halve too.
// Accumulator columns of a tile of U units (4U columns, e.g. 256 for U = 64):
//   [0,U) W_in·x   [U,2U) r = W_ir·x + W_hr·h   [2U,3U) z   [3U,4U) W_hn·h
// The x-part MMA writes columns [0,4U): its B rows are the tile's gate rows in the order [n | r | z] followed by U
// zero rows (a TMA box placed past the end of the packed matrix: out-of-bounds rows arrive as zeros), so it also CLEARS
// the W_hn·h columns; the h-part MMA then accumulates the gate rows [r | z | n] onto columns [U,4U).  One MMA per
// k-step and part — the kernel pays for every tcgen05.mma instruction, and the earlier layout (n_x + W_hn·h together,
// n_x again by a second N = U MMA per x k-step) spent 20 of its 104 MMAs per step on keeping n_x apart.
// With cta_group::2 CTA c supplies rows [c·N/2, (c+1)·N/2) of the B operand: chunk q = (N/U)·c + i of U/2 rows is
// rows [half·U/2, +U/2) of gate order[q / 2], half = q % 2 — three (x-part: + one zero) boxes per CTA and stage.
// Barriers: TMA of both CTAs counts bytes on the LEADER's full barrier (cp.async.bulk.tensor
// .cta_group::2); tcgen05.commit multicasts the stage release and "accumulator ready" to both CTAs;
// the epilogue warps of both CTAs release the accumulator on the leader's barrier (mapa + remote arrive).
//
// ROW-BLOCK INTERLEAVING (RB = 2): a step of one row block is a serial chain — publish h_t -> acquire ->
// TMA -> MMA -> gates — of which only a third is tensor work, so with one row block per pair the tensor
// core idles for two thirds of every step (ncu: tensor pipe 48 %).  With RB = 2 a pair owns TWO independent
// 256-row blocks and works on them alternately ("jobs" j = t·RB + b, accumulator j mod NACC): while block
// 0 waits for its neighbours' h_t, the tensor core runs block 1.  Same step latency, HALF the SMs — the
// other half of the device is free for question-independent work (the wide ReGAT projection / the W_v
// projection run there at the same time, api.cu `overlap`).
//
// TOKEN-TABLE FORM (TABLE = true; question encoder of the forward path): the input half of the gates depends on the
// token only, gi(v) = W_ih·emb[v] + b — a [vocabulary, 3H] table that engine.prepare_weights builds once per weight
// version (fp16, r/z/n biases folded in).  The x-part (embedding gather kernel, 24 % of the MMAs, 25 % of the operand
// bytes a CTA pulls out of L2 every step) disappears: the epilogue threads, who own the accumulator rows anyway, WRITE
// gi[token(row, t)] (and b_hn) into the accumulator of a later job with tcgen05.st while the tensor core works on the
// other accumulator — after h_t has been published, so the table reads stay out of the step-to-step chain — and the
// h-part MMAs accumulate on top.  "accumulator drained" (tempty) then also means "re-initialised".
#include <stdlib.h>
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace vqa {
namespace grup {

using namespace tc;

constexpr int PACK_UNITS = 64;                    // the packed weights come in 192-row blocks of 64 units ([r|z|n])
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 16;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;   // 640
constexpr int W_PREFETCH = 4;
constexpr int A_BYTES = BM * BK * 2;              // 16 KB
constexpr int TMEM_COLS = 512;

// UNITS = hidden units per CTA pair: 64 normally; 32 when the batch is so small that 64-unit tiles would leave
// half of the SMs idle (B <= 512 at H = 1024) — twice the CTAs, half the W tile and half the gate math per CTA.
template <int UNITS, bool TABLE = false> struct Cfg {
  static constexpr int ACC_STRIDE = 4 * UNITS;             // columns of one accumulator: [r|z|n] x 2 halves, then n_x
  static constexpr int NACC = TMEM_COLS / ACC_STRIDE;      // accumulators in flight: 2 (64 units) / 4 (32 units)
  static constexpr int HALF_UNITS = UNITS / 2;             // units whose W rows one CTA holds
  static constexpr int UPT = UNITS / (EPI_WARPS / 4);      // units per epilogue thread (16 / 8)
  static constexpr int CHUNK_BYTES = HALF_UNITS * BK * 2;  // one box of U/2 gate rows (4 KB / 2 KB)
  static constexpr int W_BYTES = (TABLE ? 3 : 4) * CHUNK_BYTES;   // this CTA's half of the widest B tile (x-part: N = 4U)
  static constexpr int STAGE_BYTES = A_BYTES + W_BYTES;    // 32 KB / 24 KB (token-table form: 28 KB / 22 KB)
  static constexpr int STAGES = TABLE ? (UNITS == 64 ? 6 : 8) : (UNITS == 64 ? 7 : 9);
  static constexpr int COL_NH = 3 * UNITS;                 // W_hn·h
  // token-table form: the 128 table rows of the next job, [r | z | n] x 32 units per 32-unit block; the row stride is an
  // ODD number of 16-byte words so that 8 consecutive rows (lanes) hit 8 different 16-byte bank groups (LDS.128)
  static constexpr int GI_ROW_BYTES = 3 * UNITS * 2;
  static constexpr int GI_STRIDE = GI_ROW_BYTES + 16;
  static constexpr int GI_BYTES = TABLE ? BM * GI_STRIDE : 0;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + GI_BYTES + 256 + 4 * UNITS * 4;
  static_assert(SMEM_BYTES <= 232448, "shared memory per CTA");
  static_assert((GI_STRIDE / 16) % 2 == 1, "row stride of the table tile");
};

struct Params {
  int B, T, H, E_pad, tiles_n, num_ctas;
  const float* bias;            // packed [4H]: b_ir+b_hr | b_iz+b_hz | b_in | b_hn
  __nv_bfloat16* h_op[2];       // bf16 state, double buffered over steps
  float* h_last;                // [B,H] f32 or NULL
  __nv_bfloat16* h_last_lp;     // [B,H] bf16 or NULL
  __nv_bfloat16* h_all;         // every state, or NULL: [B,T,H] (sequence form, see gru_tc.cu) or, time-major, [T,Bfull,H]
  int* counter;                 // per-row-block (128 rows) arrival counters, zero on entry
  // training form (train.cu): states time-major, gates saved for the backward pass (f32 [T,Bfull,H] each, chunk offset applied)
  int time_major, Bfull, b0;
  int thread_fences;            // 1 = every epilogue thread fences before the publish barrier (VQA_B200_GRU_FENCE=1)
  int debug;                    // VQA_B200_GRU_DEBUG bits, timing experiments only (results are wrong): 1 = no MMAs,
                                // 2 = no TMA loads, 4 = no wait for the other CTAs' h_t, 8 = no gate arithmetic / stores
  float *save_r, *save_z, *save_n, *save_hn, *save_h;   // save_h: slot t = state AFTER step t
  // token-table form: gi_table fp16 [ntoken_rows, H/32, 3, 32] (per 32-unit block the gates r|z|n; biases b_ir+b_hr |
  // b_iz+b_hz | b_in folded in)
  const __half* gi_table;
  const int64_t* tokens;        // [B,T]
  int ntoken_rows;
  unsigned gi_pause_ns;         // pause between the loader warp's bursts of 32 row copies (VQA_B200_GRU_GI_PAUSE, ns)
};

__device__ __forceinline__ void half8_to_f32(const uint4& v, uint32_t (&o)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    o[2 * i] = __float_as_uint(f.x); o[2 * i + 1] = __float_as_uint(f.y);
  }
}

__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }

template <int UNITS, int RB, bool TABLE>
__global__ void __launch_bounds__(THREADS, 1)
gru_pair_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmH0,
                const __grid_constant__ CUtensorMap tmH1, const __grid_constant__ CUtensorMap tmWx,
                const __grid_constant__ CUtensorMap tmWh, const Params p) {
  using C = Cfg<UNITS, TABLE>;
  constexpr int HALF_UNITS = C::HALF_UNITS, UPT = C::UPT, STAGE_BYTES = C::STAGE_BYTES, STAGES = C::STAGES, COL_NH = C::COL_NH;
  constexpr int CHUNK_BYTES = C::CHUNK_BYTES;
  constexpr int ACC_STRIDE = C::ACC_STRIDE, NACC = C::NACC;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + NACC + a); };
  static_assert(8 * (2 * STAGES + 2 * NACC + 3) <= 256, "barrier block");
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 2 * NACC);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 2 * NACC));
  const uint32_t gi_full = bars + 8u * (2 * STAGES + 2 * NACC + 1), gi_empty = gi_full + 8u;   // token-table tile
  float* bias_s = reinterpret_cast<float*>(base_ptr + STAGES * STAGE_BYTES + 256);   // [4][64]
  const uint32_t gi_s = bars + 256u + 4u * UNITS * 4u;                                // [128 rows][GI_STRIDE] (TABLE)
  const uint8_t* gi_ptr = base_ptr + STAGES * STAGE_BYTES + 256 + 4 * UNITS * 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();          // 0 = leader of the pair
  const bool lead = crank == 0;
  const int pid = blockIdx.x >> 1;
  const int grp = pid / p.tiles_n, n_blk = pid % p.tiles_n;
  // the pair owns RB blocks of 256 rows; this CTA's 128 rows of block b are row block m_blk_of(b)
  auto m_blk_of = [&](int b) { return 2 * (grp * RB + b) + (int)crank; };
  const int u0 = n_blk * UNITS;
  const int kb_x = p.E_pad / BK, kb_h = p.H / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmH0); tma_prefetch_desc(&tmH1);
    tma_prefetch_desc(&tmWx); tma_prefetch_desc(&tmWh);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    for (int a = 0; a < NACC; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 2 * EPI_WARPS); }
    mbar_init(gi_full, 1); mbar_init(gi_empty, EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2cta(tmem_slot, TMEM_COLS);
  if (warp == 3) {
    for (int i = lane; i < 4 * UNITS; i += 32) bias_s[i] = p.bias[(i / UNITS) * p.H + u0 + (i % UNITS)];
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                // the peer signals / fills through these barriers
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();

  // this CTA's half of the B tile: chunks q = n_chunks·crank + i of HALF_UNITS rows, chunk q = rows [half·HALF, +HALF) of
  // gate order[q / 2] of the tile's units (packed matrix viewed as [3H/64 gate blocks r,z,n][64 rows][K]); x-part: gate
  // order n, r, z and a fourth, all-zero chunk per CTA pair half (out-of-bounds box); h-part: r, z, n.  The bytes the
  // stage's barrier expects are the CTA pair's: x_part ? 2·(A + 4 chunks) : 2·(A + 3 chunks).
  const int gate_blk0 = (u0 / PACK_UNITS) * 3, unit_row0 = u0 % PACK_UNITS, oob_blk = 3 * p.H / PACK_UNITS;
  auto load_w = [&](uint32_t sw, const CUtensorMap* map, uint32_t bar, int col, bool x_part) {
    const int n_chunks = x_part ? 4 : 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= n_chunks) break;
      const int q = n_chunks * (int)crank + i;             // 0..7 (x) / 0..5 (h) over the pair
      const int pos = q >> 1, half = q & 1;
      const int gate = x_part ? (pos == 0 ? 2 : pos - 1) : pos;          // x: n, r, z, (zeros)   h: r, z, n
      const int blk = (x_part && pos == 3) ? oob_blk : gate_blk0 + gate;
      tma_load_3d_2cta(sw + i * CHUNK_BYTES, map, bar, col, unit_row0 + half * HALF_UNITS, blk);
    }
  };
  constexpr uint32_t X_STAGE_TX = 2 * (A_BYTES + 4 * CHUNK_BYTES), H_STAGE_TX = 2 * (A_BYTES + 3 * CHUNK_BYTES);

  if (warp == 0) {
    // ===== TMA producer (both CTAs; bytes are counted on the leader's full barrier) =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const bool noload = (p.debug & 2) != 0;
      for (int t = 0; t < p.T; ++t) {
       for (int b = 0; b < RB; ++b) {
        const int m_blk = m_blk_of(b), m0 = m_blk * BM;
        for (int kb = 0; kb < (TABLE ? 0 : kb_x); ++kb) {         // x-part: no dependence on h
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
          if (lead) mbar_arrive_expect_tx(full_bar(stage), noload ? 0u : X_STAGE_TX);
          if (!noload) {
            tma_load_2d_2cta(sa, &tmX, full_bar(stage), t * p.E_pad + kb * BK, m0);
            load_w(sw, &tmWx, full_bar(stage), kb * BK, true);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (t > 0) {
          const int npre = kb_h < W_PREFETCH ? kb_h : W_PREFETCH;
          int pre_stage[W_PREFETCH];
          for (int kb = 0; kb < npre; ++kb) {                     // W_h tiles do not depend on h: start them early
            mbar_wait(empty_bar(stage), phase ^ 1);
            if (lead) mbar_arrive_expect_tx(full_bar(stage), noload ? 0u : H_STAGE_TX);
            if (!noload) load_w(base + stage * STAGE_BYTES + A_BYTES, &tmWh, full_bar(stage), kb * BK, false);
            pre_stage[kb] = stage;
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          // h_{t-1}[rows of this CTA, :] is published by the tiles_n CTAs that own these rows
          const int target = t * p.tiles_n;
          if (!(p.debug & 4)) while (ld_acquire_gpu(p.counter + m_blk) < target) { }
          fence_proxy_async_all();
          const CUtensorMap* tmH = ((t - 1) & 1) ? &tmH1 : &tmH0;
          const int hcol = (p.h_all && !p.time_major) ? (t - 1) * p.H : 0;
          const int hrow = p.time_major ? (t - 1) * p.Bfull + p.b0 + m0 : m0;
          for (int kb = 0; kb < npre; ++kb)
            if (!noload) tma_load_2d_2cta(base + pre_stage[kb] * STAGE_BYTES, tmH, full_bar(pre_stage[kb]), hcol + kb * BK, hrow);
          for (int kb = npre; kb < kb_h; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1);
            const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
            if (lead) mbar_arrive_expect_tx(full_bar(stage), noload ? 0u : H_STAGE_TX);
            if (!noload) {
              tma_load_2d_2cta(sa, tmH, full_bar(stage), hcol + kb * BK, hrow);
              load_w(sw, &tmWh, full_bar(stage), kb * BK, false);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
       }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread of the LEADER CTA drives both tensor cores =====
    if (lane == 0 && lead) {
      constexpr uint32_t idesc_x = make_idesc_bf16(2 * BM, 4 * UNITS);     // M = 256, N = 256 (128): [n | r | z | zeros]
      constexpr uint32_t idesc_h = make_idesc_bf16(2 * BM, 3 * UNITS);     // M = 256, N = 192 (96):  [r | z | n]
      int stage = 0; uint32_t phase = 0;
      uint32_t job = 0;                                  // j = t·RB + b: accumulator j mod NACC, its (j / NACC)-th use
      for (int t = 0; t < p.T; ++t) {
       for (int b = 0; b < RB; ++b, ++job) {
        const uint32_t acc = job % NACC;
        const uint32_t acc_phase = (job / NACC) & 1u;
        // TABLE: every use of an accumulator, the first one too, waits for the epilogue threads' initialisation
        mbar_wait(tempty_bar(acc), TABLE ? acc_phase : (acc_phase ^ 1));
        tcgen05_fence_after();
        const uint32_t d = tmem_base + acc * ACC_STRIDE;
        for (int kb = 0; kb < (TABLE ? 0 : kb_x); ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
          const uint64_t adesc = make_sw128_kmajor_desc(sa), wdesc = make_sw128_kmajor_desc(sw);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)                  // columns [0,4U): the first one also clears W_hn·h
            if (!(p.debug & 1)) umma_bf16_2cta(d, adesc + 2 * k, wdesc + 2 * k, idesc_x, (kb | k) != 0);
          umma_commit_2cta(empty_bar(stage), 0b11);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (t > 0) {
          for (int kb = 0; kb < kb_h; ++kb) {
            mbar_wait(full_bar(stage), phase);
            tcgen05_fence_after();
            const uint32_t sa = base + stage * STAGE_BYTES, sw = sa + A_BYTES;
            const uint64_t adesc = make_sw128_kmajor_desc(sa), wdesc = make_sw128_kmajor_desc(sw);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              if (!(p.debug & 1)) umma_bf16_2cta(d + UNITS, adesc + 2 * k, wdesc + 2 * k, idesc_h, 1u);
            umma_commit_2cta(empty_bar(stage), 0b11);
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
        umma_commit_2cta(tfull_bar(acc), 0b11);
       }
      }
    }
  } else if (TABLE && warp == 3) {
    // ===== token-table loader: per job, this CTA's 128 rows x [r|z|n] x UNITS of the table as 128 bulk copies =====
    // (per-thread global loads of these rows — 32 different sectors per warp instruction — kept the SM's load/store
    // pipeline busy for ~1.5 us per step right when the next step's acquire / TMA issue / mbarrier polls needed it)
    const uint32_t total_jobs = (uint32_t)p.T * RB;
    for (uint32_t jn = 0; jn < total_jobs; ++jn) {
      const int tn = (int)(jn / RB), bn = (int)(jn % RB);
      mbar_wait(gi_empty, (jn & 1u) ^ 1u);                           // the epilogue warps have read the previous tile
      if (lane == 0) mbar_arrive_expect_tx(gi_full, (p.debug & 16) ? 0u : (uint32_t)(BM * C::GI_ROW_BYTES));
      __syncwarp();
      if (p.debug & 16) continue;                                    // debug 16: no table reads
#pragma unroll
      for (int i = 0; i < BM / 32; ++i) {
        // the copies are needed a whole step from now: issue them in 4 bursts of 32 with pauses in between, so that the
        // TMA unit takes the next step's h / W tile loads (the critical path) between the bursts instead of behind all 128
        if (i > 0 && p.gi_pause_ns > 0) __nanosleep(p.gi_pause_ns);
        const int r = i * 32 + lane;
        int rown = m_blk_of(bn) * BM + r;
        rown = rown < p.B ? rown : p.B - 1;                          // rows beyond the batch: any valid row (masked later)
        long long tok = (long long)p.tokens[(size_t)rown * p.T + tn];
        tok = tok < 0 ? 0 : (tok >= p.ntoken_rows ? p.ntoken_rows - 1 : tok);
        const __half* src = p.gi_table + (size_t)tok * 3 * p.H + (size_t)(u0 / 32) * 96;
        bulk_g2s(gi_s + (uint32_t)r * C::GI_STRIDE, src, (uint32_t)C::GI_ROW_BYTES, gi_full);
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== gate epilogue (both CTAs, own 128 rows of every block): thread = (batch row, 16 units), fp32 state in registers =====
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int uh = (warp - EPI_WARP0) >> 2;          // which 16-unit slice of the 64 units
    const int et = threadIdx.x - EPI_WARP0 * 32;
    const int ub = uh * UPT;                         // first unit (within the tile) of this thread = its W_in·x column
    float h[RB][UPT];
#pragma unroll
    for (int b = 0; b < RB; ++b)
#pragma unroll
      for (int j = 0; j < UPT; ++j) h[b][j] = 0.f;
    // bias of gate row i of the tile; the token-table form has them in the accumulator already
    auto bs = [&](int i) { return TABLE ? 0.f : bias_s[i]; };
    // TABLE: the accumulator of job jn (step jn / RB, block jn % RB) starts as gi[token(row, step)] (+ biases) and b_hn;
    // the loader warp has staged the 128 rows in shared memory (row r at gi_s + r·GI_STRIDE: per 32-unit block [r|z|n])
    const uint32_t total_jobs = (uint32_t)p.T * RB;
    auto init_acc = [&](uint32_t jn) {
      const uint32_t tr = tmem_base + (jn % NACC) * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      mbar_wait(gi_full, jn & 1u);
      const uint8_t* rowp = gi_ptr + (q * 32 + lane) * C::GI_STRIDE + (ub / 32) * 192 + (ub % 32) * 2;
      uint4 gr[UPT / 8], gz[UPT / 8], gn[UPT / 8];
#pragma unroll
      for (int c = 0; c < UPT / 8; ++c) {
        gr[c] = reinterpret_cast<const uint4*>(rowp)[c];
        gz[c] = reinterpret_cast<const uint4*>(rowp + 64)[c];
        gn[c] = reinterpret_cast<const uint4*>(rowp + 128)[c];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(gi_empty);                           // tile read: the loader may fetch the next job's
#pragma unroll
      for (int c = 0; c < UPT; c += 8) {
        uint32_t v[8];
        half8_to_f32(gn[c / 8], v); tmem_st_32x8(tr + ub + c, v);
        half8_to_f32(gr[c / 8], v); tmem_st_32x8(tr + UNITS + ub + c, v);
        half8_to_f32(gz[c / 8], v); tmem_st_32x8(tr + 2 * UNITS + ub + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __float_as_uint(bias_s[3 * UNITS + ub + c + j]);
        tmem_st_32x8(tr + COL_NH + ub + c, v);
      }
      tmem_st_wait();
    };
    if (TABLE) {
      for (uint32_t jn = 0; jn < (uint32_t)NACC && jn < total_jobs; ++jn) {
        init_acc(jn);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(jn), 0);
      }
    }
    uint32_t job = 0;
    for (int t = 0; t < p.T; ++t) {
      const bool last = (t == p.T - 1);
      const size_t slot = (size_t)t * p.Bfull * p.H;             // time-major slot of this step
      __nv_bfloat16* hdst = p.h_all ? p.h_all + (p.time_major ? slot : (size_t)t * p.H)
                                    : ((last && p.h_last_lp) ? p.h_last_lp : p.h_op[t & 1]);
      const size_t h_ld = (p.h_all && !p.time_major) ? (size_t)p.T * p.H : (size_t)p.H;
#pragma unroll
      for (int b = 0; b < RB; ++b, ++job) {
      const int m_blk = m_blk_of(b);
      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.B;
      const uint32_t acc = job % NACC;
      const uint32_t acc_phase = (job / NACC) & 1u;
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      const uint32_t trow = tmem_base + acc * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < UPT; c += 8) {
        if (p.debug & 8) break;
        uint32_t vr[8], vz[8], vni[8], vnh[8];
        tmem_ld_32x8(trow + UNITS + ub + c, vr);
        tmem_ld_32x8(trow + 2 * UNITS + ub + c, vz);
        tmem_ld_32x8(trow + COL_NH + ub + c, vnh);
        tmem_ld_32x8(trow + ub + c, vni);
        tmem_ld_wait();
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float pr = __uint_as_float(vr[j]) + bs(ub + c + j), pz = __uint_as_float(vz[j]) + bs(UNITS + ub + c + j);
          const float r = sigmoid_fast(pr), z = sigmoid_fast(pz);
          const float nh = __uint_as_float(vnh[j]) + bs(3 * UNITS + ub + c + j);
          const float pn = __uint_as_float(vni[j]) + bs(2 * UNITS + ub + c + j) + r * nh;
          const float n = tanh_fast(pn);
          const float hn = (1.f - z) * n + z * h[b][c + j];
          h[b][c + j] = hn;
          o[j] = hn;
        }
        if (row_ok) {
          uint4 w0;
          w0.x = pack_bf16x2(o[0], o[1]); w0.y = pack_bf16x2(o[2], o[3]);
          w0.z = pack_bf16x2(o[4], o[5]); w0.w = pack_bf16x2(o[6], o[7]);
          *reinterpret_cast<uint4*>(hdst + (size_t)row * h_ld + u0 + ub + c) = w0;
        }
      }
      const bool saving = !TABLE && p.save_r != nullptr;
      if (!saving && !TABLE) {
        // accumulator buffer drained: one arrive per warp on the LEADER's barrier (2 CTAs x 16 warps)
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
      }
      if (!last) {
        // publish h_t: the state stores of all epilogue threads are ordered before the barrier; the release at gpu scope of
        // the ONE thread that then bumps the counter is cumulative over what it has observed through the barrier (PTX
        // memory model), so the other 511 threads need no fence of their own (each cost a round trip to L2 in the chain)
        if (p.thread_fences) __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        if (et == 0) {
          __threadfence();
          fence_proxy_async_all();
          red_release_gpu_add(p.counter + m_blk, 1);
        }
      }
      if (TABLE) {
        // h_t is on its way to the other CTAs: now (off the chain) prepare this accumulator for its next job and hand it
        // back to the MMA thread
        if (job + NACC < total_jobs) init_acc(job + NACC);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
      }
      if (saving) {
        // Training form: what the backward pass needs (train.cu) is written AFTER h_t has been published, so the
        // 20 extra 16-byte stores per thread stay out of the step-to-step dependency chain; the gates are
        // recomputed from the accumulators, which are released only now (the block's next job uses another buffer
        // or, with NACC == RB, waits for this release).
#pragma unroll
        for (int c = 0; c < UPT; c += 8) {
          uint32_t vr[8], vz[8], vni[8], vnh[8];
          tmem_ld_32x8(trow + UNITS + ub + c, vr);
          tmem_ld_32x8(trow + 2 * UNITS + ub + c, vz);
          tmem_ld_32x8(trow + COL_NH + ub + c, vnh);
          tmem_ld_32x8(trow + ub + c, vni);
          tmem_ld_wait();
          float gr[8], gz[8], gn[8], ghn[8], hs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            gr[j] = sigmoid_fast(__uint_as_float(vr[j]) + bias_s[ub + c + j]);
            gz[j] = sigmoid_fast(__uint_as_float(vz[j]) + bias_s[UNITS + ub + c + j]);
            ghn[j] = __uint_as_float(vnh[j]) + bias_s[3 * UNITS + ub + c + j];
            gn[j] = tanh_fast(__uint_as_float(vni[j]) + bias_s[2 * UNITS + ub + c + j] + gr[j] * ghn[j]);
            hs[j] = h[b][c + j];
          }
          if (row_ok) {
            const size_t off = slot + (size_t)row * p.H + u0 + ub + c;
            store8(p.save_r + off, gr); store8(p.save_z + off, gz); store8(p.save_n + off, gn); store8(p.save_hn + off, ghn);
            store8(p.save_h + off, hs);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
      }
      }
    }
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int row = m_blk_of(b) * BM + q * 32 + lane;
      if (row < p.B && p.h_last) {
        float* dst = p.h_last + (size_t)row * p.H + u0 + ub;
#pragma unroll
        for (int j = 0; j < UPT; j += 4)
          *reinterpret_cast<float4*>(dst + j) = make_float4(h[b][j], h[b][j + 1], h[b][j + 2], h[b][j + 3]);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();                                // no CTA leaves while its peer may still use its memory / barriers
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS);
}

}  // namespace grup

// Same contract as gru_persistent (gru_tc.cu).  With `save` (training form): h_all is time-major [T,B,H] (slot t = state
// after step t) and the gates r, z, n, W_hn·h + b_hn and the f32 states are stored per step for the backward pass.  Returns VQA_ERR_UNSUPPORTED (without setting the error text as a
// failure of the call) when the pair launch is not possible; the caller then uses the single-CTA kernel.
// sm_limit > 0: use at most that many SMs (the caller runs something else on the rest at the same time).
template <int UNITS, int RB, bool TABLE>
static int gru_pair_t(const void* X, int B, int T, int H, int E_pad, const void* wx_p, const void* wh_p,
                      const float* bias_p, void* h_op, int* counter, float* h_last, void* h_last_lp, void* h_all,
                      const GruTrainSave* save, int sm_limit, const GruTokenTable* tab, cudaStream_t s) {
  using namespace grup;
  using C = Cfg<UNITS, TABLE>;
  constexpr int HALF_UNITS = C::HALF_UNITS, SMEM_BYTES = C::SMEM_BYTES;
  auto kernel = gru_pair_kernel<UNITS, RB, TABLE>;
  if (H % PACK_UNITS != 0 || E_pad % tc::BK != 0) return VQA_ERR_UNSUPPORTED;
  const int tiles_n = H / UNITS;
  const int dev = current_device();
  static DeviceInt usable_dev;                       // number of pairs the device can hold at once (0 = pairs unusable)
  int& usable = usable_dev.at(dev);
  if (usable < 0) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess) {
      (void)cudaGetLastError();
      usable = 0;
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (sm_count() / 2)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int clusters = 0;
      if (cudaOccupancyMaxActiveClusters(&clusters, (const void*)kernel, &cfg) != cudaSuccess) { (void)cudaGetLastError(); clusters = 0; }
      usable = clusters > 0 ? clusters : 0;
    }
  }
  int pairs_max = usable;
  if (sm_limit > 0 && sm_limit / 2 < pairs_max) pairs_max = sm_limit / 2;
  const int max_groups = pairs_max / tiles_n;        // groups of RB x 256 rows that can be co-resident
  if (max_groups < 1) return VQA_ERR_UNSUPPORTED;
  const int rows_per_group = 2 * RB * tc::BM;
  // cooperative launch = the runtime checks co-residency of the whole grid; VQA_B200_GRU_COOP=0 launches plainly (the
  // grid is sized to fit either way)
  static int coop = -1;
  if (coop < 0) { const char* e = getenv("VQA_B200_GRU_COOP"); coop = (e && e[0] == '0') ? 0 : 1; }
  CUtensorMap tmWx, tmWh;
  int rc;
  {
    // packed W [3H, K] viewed as [3H/64 gate blocks][64 rows][K]: box = {64 cols, HALF_UNITS rows of one gate block}
    const int box[3] = {tc::BK, HALF_UNITS, 1};
    const long long dx[3] = {E_pad, PACK_UNITS, 3LL * H / PACK_UNITS}, sx[2] = {2LL * E_pad, 2LL * E_pad * PACK_UNITS};
    const long long dh[3] = {H, PACK_UNITS, 3LL * H / PACK_UNITS}, sh[2] = {2LL * H, 2LL * H * PACK_UNITS};
    if ((rc = tc::make_tensor_map_bf16_nd(&tmWh, wh_p, 3, dh, sh, box))) return rc;
    if (TABLE) tmWx = tmWh;                            // never used
    else if ((rc = tc::make_tensor_map_bf16_nd(&tmWx, wx_p, 3, dx, sx, box))) return rc;
  }
  for (int b0 = 0; b0 < B; b0 += max_groups * rows_per_group) {
    const int Bc = (B - b0 < max_groups * rows_per_group) ? B - b0 : max_groups * rows_per_group;
    const int groups = (Bc + rows_per_group - 1) / rows_per_group;   // padded to whole groups; extra rows are masked
    const int tiles_m = groups * 2 * RB;
    const __nv_bfloat16* Xc = (const __nv_bfloat16*)X + (size_t)b0 * T * E_pad;
    __nv_bfloat16* h0 = (__nv_bfloat16*)h_op + (size_t)b0 * H;
    __nv_bfloat16* h1 = (__nv_bfloat16*)h_op + (size_t)B * H + (size_t)b0 * H;
    CUtensorMap tmX, tmH0, tmH1;
    if (!TABLE && (rc = tc::make_tensor_map_bf16(&tmX, Xc, Bc, (long long)T * E_pad, (long long)T * E_pad, tc::BM))) return rc;
    const bool tmajor = save != nullptr;
    __nv_bfloat16* hall = h_all ? (__nv_bfloat16*)h_all + (tmajor ? (size_t)b0 * H : (size_t)b0 * T * H) : nullptr;
    if (hall && tmajor) {                               // [T*B rows, H]: step t reads rows (t-1)*B + b0 + m0 ..
      if ((rc = tc::make_tensor_map_bf16(&tmH0, h_all, (long long)T * B, H, H, tc::BM))) return rc;
      tmH1 = tmH0;
    } else if (hall) {
      if ((rc = tc::make_tensor_map_bf16(&tmH0, hall, Bc, (long long)T * H, (long long)T * H, tc::BM))) return rc;
      tmH1 = tmH0;
    } else {
      if ((rc = tc::make_tensor_map_bf16(&tmH0, h0, Bc, H, H, tc::BM))) return rc;
      if ((rc = tc::make_tensor_map_bf16(&tmH1, h1, Bc, H, H, tc::BM))) return rc;
    }
    if (TABLE) tmX = tmH0;                             // never used
    Params p;
    p.gi_table = TABLE ? (const __half*)tab->gi_table : nullptr;
    p.tokens = TABLE ? tab->tokens + (size_t)b0 * T : nullptr;
    p.ntoken_rows = TABLE ? tab->ntoken_rows : 0;
    static int pause = -1;
    if (pause < 0) { const char* e = getenv("VQA_B200_GRU_GI_PAUSE"); pause = e ? atoi(e) : 0; }
    p.gi_pause_ns = (unsigned)pause;
    p.B = Bc; p.T = T; p.H = H; p.E_pad = E_pad; p.tiles_n = tiles_n; p.num_ctas = 2 * groups * tiles_n;
    p.bias = bias_p; p.h_op[0] = h0; p.h_op[1] = h1;
    p.h_last = h_last ? h_last + (size_t)b0 * H : nullptr;
    p.h_all = hall;
    p.h_last_lp = h_last_lp ? (__nv_bfloat16*)h_last_lp + (size_t)b0 * H : nullptr;
    p.counter = counter;
    p.time_major = tmajor ? 1 : 0; p.Bfull = B; p.b0 = b0;
    static int fences = -1;
    if (fences < 0) { const char* e = getenv("VQA_B200_GRU_FENCE"); fences = (e && e[0] == '1') ? 1 : 0; }
    p.thread_fences = fences;
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("VQA_B200_GRU_DEBUG"); dbg = e ? atoi(e) : 0; }
    p.debug = dbg;
    const size_t so = (size_t)b0 * H;
    p.save_r = tmajor ? save->R + so : nullptr; p.save_z = tmajor ? save->Z + so : nullptr;
    p.save_n = tmajor ? save->N + so : nullptr; p.save_hn = tmajor ? save->HN + so : nullptr;
    p.save_h = tmajor ? save->Hs + so : nullptr;
    if (tiles_m > 64) return fail(VQA_ERR_INVALID, "gru_pair: %d row blocks per launch exceed the counter block", tiles_m);
    // the counters; the first launch of a token-table call also clears what the caller parked right behind the 256-byte
    // counter block (the gather kernel that used to do it is gone)
    const size_t zero_bytes = (TABLE && b0 == 0 && tab->zero_after_counter) ? 256 + tab->zero_after_counter : sizeof(int) * tiles_m;
    VQA_CUDA_CHECK(cudaMemsetAsync(counter, 0, zero_bytes, s));
    void* args[] = {(void*)&tmX, (void*)&tmH0, (void*)&tmH1, (void*)&tmWx, (void*)&tmWh, (void*)&p};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(p.num_ctas); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = coop ? 2 : 1;
    VQA_CUDA_CHECK(cudaLaunchKernelExC(&cfg, (const void*)kernel, args));
    count_launch();
  }
  return VQA_OK;
}

// Tile choice.  One 256-row block per pair (RB = 1): 64-unit tiles, or 32-unit tiles when 64-unit tiles would use at most
// half of the SMs (B <= 512 at H = 1024).  Two interleaved row blocks per pair (RB = 2, see the header): when the caller
// gives the GRU only a share of the device (sm_limit > 0: the other SMs run question-independent GEMM tiles at the same
// time) and the batch has at least two 256-row blocks.  VQA_B200_GRU_CFG = "64x1" | "32x1" | "64x2" | "32x2" forces one.
int gru_pair(const void* X, int B, int T, int H, int E_pad, const void* wx_p, const void* wh_p,
             const float* bias_p, void* h_op, int* counter, float* h_last, void* h_last_lp, void* h_all,
             const GruTrainSave* save, int sm_limit, const GruTokenTable* tab, cudaStream_t s) {
  if (tab && (save || h_all || !tab->gi_table || !tab->tokens)) return fail(VQA_ERR_INVALID, "gru_pair: the token-table form has no sequence / training variant");
  static int forced = -1;                            // 0 = automatic, else 10 * units + rb
  if (forced < 0) {
    forced = 0;
    const char* e = getenv("VQA_B200_GRU_CFG");
    if (e) {
      int u = 0, r = 0;
      if (sscanf(e, "%dx%d", &u, &r) == 2 && (u == 32 || u == 64) && (r == 1 || r == 2)) forced = 10 * u + r;
    }
    const char* e64 = getenv("VQA_B200_GRU_UNITS");  // older switch: 64 forces the wide tile
    if (!forced && e64 && atoi(e64) == 64) forced = 641;
  }
  if (H % grup::PACK_UNITS != 0) return VQA_ERR_UNSUPPORTED;
  const int sms = sm_count();
  const int budget = (sm_limit > 0 && sm_limit < sms) ? sm_limit : sms;
  const int row_pairs = (B + 2 * tc::BM - 1) / (2 * tc::BM);
  auto ctas = [&](int units, int rb) { return 2 * ((row_pairs + rb - 1) / rb) * (H / units); };
  int cfg = forced;
  if (!cfg) {
    if (2 * ctas(64, 1) <= budget) cfg = 321;                 // small batch: narrower tiles fill the SMs
    else if (ctas(64, 1) <= budget) cfg = 641;                // one row block per pair fits: lowest latency
    else if (!save && row_pairs >= 2 && ctas(64, 2) <= budget) cfg = 642;   // half the SMs: two interleaved row blocks
    else cfg = 641;                                           // several launches
  }
  // (the token-table form runs on every tile shape: a 128-question shard of an 8-GPU job takes 32-unit tiles and must
  // reproduce its rows of the unsharded batch bit for bit, so the arithmetic may not depend on the tile choice)
#define VQA_GRU_CALL(U, R) (tab ? gru_pair_t<U, R, true>(X, B, T, H, E_pad, wx_p, wh_p, bias_p, h_op, counter, h_last, h_last_lp, h_all, save, sm_limit, tab, s) \
                                : gru_pair_t<U, R, false>(X, B, T, H, E_pad, wx_p, wh_p, bias_p, h_op, counter, h_last, h_last_lp, h_all, save, sm_limit, tab, s))
  switch (cfg) {
    case 321: return VQA_GRU_CALL(32, 1);
    case 322: return VQA_GRU_CALL(32, 2);
    case 642: return VQA_GRU_CALL(64, 2);
    default: return VQA_GRU_CALL(64, 1);
  }
#undef VQA_GRU_CALL
}

}  // namespace vqa
