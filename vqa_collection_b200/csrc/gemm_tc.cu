// tcgen05 / TMEM / TMA implementation of vqa_linear for bf16 operands (sm_100a).
//
// Reference arithmetic: modules/modules.py:13-60 (FCNet = weight_norm(nn.Linear)
// + ReLU), attention.py:70-75 (W_v projection ⊙ W_q projection → 1-wide logit
// layer), gcn.py:101-103 and modules.py:92-93 (plain nn.Linear maps).
//
// Design (B200-first, not a translation of anything in the reference — it has no
// kernels): persistent warp-specialised CTA per SM.
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D loads of a 128x64 A box and a
//             BNx64 W box (bf16, 128-byte swizzle) into a STAGES-deep smem ring,
//             completion on mbarriers (complete_tx)
//   warp 1    MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//             (M=128, N=BN, K=16) with smem descriptors; fp32 accumulators live in
//             TMEM, double buffered (2 x BN columns) so the epilogue of tile i
//             overlaps the main loop of tile i+1; tcgen05.commit frees smem slots
//   warp 2    TMEM allocator (tcgen05.alloc / dealloc)
//   warps 4-7 epilogue: tcgen05.ld 32x32b (thread = accumulator row), fused
//             scale·acc + bias → ReLU → ⊙mul → {store | row-reduction with logit_w}
// A is [M,K] row-major, W is [N,K] row-major (nn.Linear layout): both operands are
// K-major, so no transposes are ever materialised.  The backward GEMMs of the training step
// (dX = dY·W, dW = dYᵀ·X) read the SAME row-major tensors with the contraction index as the
// row index: `trans_a` / `trans_w` select MN-major operand tiles (TMA boxes {64 mn, 64 k-rows},
// UMMA descriptors with the M/N-major bit) — again no transpose kernels.
#include <stdlib.h>

#include "tc_common.cuh"

namespace vqa {

namespace tc {

constexpr int THREADS = 256;
constexpr int EPI_WARP0 = 4;           // warps 4..7 (warp_id % 4 selects the TMEM lane quarter)

// PAIR: two CTAs on one TPC work on a 256 x BN tile with tcgen05.mma.cta_group::2 — each CTA loads its own 128 rows of A
// and HALF of the W tile (the tensor cores of both SMs read both halves), so the shared-memory fill and the B-operand
// reads per SM shrink by a third / a half; one thread of the leader CTA issues the MMAs (see gru_pair.cu for the protocol).
// SPLIT (dtype VQA_F16X2, the fp32-class mode): every operand is a pair of fp16 planes, x = hi + lo'·2^-11 with
// hi = fp16(x), lo' = fp16((x - hi)·2^11) (22 significant bits, no subnormal loss of the residual: Ootomo & Yokota,
// "Recovering single precision accuracy from Tensor Cores while surpassing the FP32 theoretical peak performance").  A stage
// holds [A_hi | A_lo | W_hi | W_lo]; per k-step THREE MMAs: hi·hi into the main accumulator, hi·lo' and lo'·hi into a second one
// that the epilogue adds scaled by 2^-11 (the lo'·lo' term, 2^-22 relative, is dropped).  A tile therefore owns 2·BN TMEM
// columns, and the accumulators are double buffered only where 4·BN columns fit (BN <= 128).
template <int BN, bool PAIR = false, bool SPLIT = false> struct Cfg {
  static constexpr int PLANES = SPLIT ? 2 : 1;
  static constexpr int A_BYTES = BM * BK * 2;                // one plane
  static constexpr int B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2;
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  static constexpr int STAGES = SPLIT ? (PAIR ? (BN >= 192 ? 3 : (BN >= 96 ? 4 : 5)) : (BN >= 192 ? 2 : (BN >= 128 ? 3 : 4)))
                                      : (PAIR ? 6 : (BN >= 256 ? 4 : (BN >= 192 ? 5 : (BN >= 128 ? 6 : 8))));
  static constexpr int TILE_COLS = PLANES * BN;              // TMEM columns of one tile: [main | correction]
  static constexpr int NACC = (2 * TILE_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_COLS = pow2_ge(NACC * TILE_COLS);
  static constexpr int ACC_STRIDE = TMEM_COLS / NACC;
  static constexpr int PARAM_FLOATS = 2 * 3 * BN;            // 2 acc stages x (scale,bias,logit_w)
  static constexpr int STAGE_OUT_FLOATS = 4 * 32 * 33;       // per epilogue warp: 32 rows x 32 cols (+1 pad) f32
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + 256 /*barriers*/ +
                                    PARAM_FLOATS * 4 + STAGE_OUT_FLOATS * 4;
};

struct Params {
  int M, N, K;
  const float* scale; const float* bias; int relu;
  const float* mul; int ld_mul; int mul_row_div;
  const float* add; int ld_add; int add_row_div;
  const void* mask; int ld_mask; int mask_bf16;
  const float* logit_w;
  void* out; int ldo; int out_bf16; int n_parts;
  int out_split; size_t out_plane;            // fp16 plane pair: lo' plane out_plane bytes behind the hi plane
  // GRU step epilogue (template GRU; fp32-class mode): the tile's 192 columns are [r | z | n] x 64 units of h·W_hhᵀ
  // (gate-interleaved packed W_hh); gi = gru_table[token(row, t)] (f32 [rows, 3H], b_ih folded in), torch gate order
  // One launch runs steps [gru_t, gru_t_end): step t reads the state's plane pair from buffer (t-1)&1 of the 3-D map tmA
  // ([2 buffers x 2 planes][B][H]) and writes buffer t&1 (the last step of the sequence: gru_planes_last / gru_hout_last
  // when given); the f32 state gru_h is updated in place.  With more than one step per launch every CTA is co-resident and
  // a finished tile bumps gru_counter[its 128-row block] (zero on entry); a step's loads wait until all tiles of the
  // previous step of their row block have arrived (release / acquire + fence.proxy.async, as in gru_pair.cu).
  const long long* gru_tokens; const float* gru_table; const float* gru_bhh;
  float* gru_h; float* gru_hout_last; void* gru_planes; void* gru_planes_last; int* gru_counter;
  int gru_T, gru_t, gru_t_end, gru_H, gru_rows;
  int gru_debug;   // VQA_B200_GRUS_DEBUG, timing experiments only (results wrong): 1 = no MMAs, 2 = no wait for h_{t-1},
                   // 4 = no gate arithmetic / table reads / stores, 8 = no TMA loads
  int tiles_m, tiles_n;
  int tile_begin, tile_end;                  // this launch walks tiles [tile_begin, tile_end) of the tile order
  float leaky_slope; int add_after_act; int sigmoid;
  // fused lowest-index argmax over n (wrapper.py:14): keys u64 [M] (value order bits << 32 | ~n), per-row-block tile
  // counters [tiles_m] — both zero on entry — and the int64 labels written by the LAST tile of each row block
  unsigned long long* amax_keys; int* amax_cnt; long long* amax_label;
  // optional progress counters [tiles_m], zero on entry: every finished tile adds 1 to the counter of its 128-row block
  // (release), so a consumer kernel running BESIDE this one can start on a row block as soon as all tiles_n tiles of it
  // are in memory (graph_attn_tc.cu chases the wide projection this way)
  int* progress;
};

// ---- the kernel ----------------------------------------------------------------
// GRU step form: 8 epilogue warps (the gate update with its table gathers and precise exp / tanh is 3-4x the work of a
// linear epilogue): warp w handles TMEM lane quarter w % 4 and units [32·((w-4)/4), +32) of the tile
constexpr int GRU_EPI_WARPS = 8;
constexpr int GRU_THREADS = EPI_WARP0 * 32 + GRU_EPI_WARPS * 32;

template <int BN, bool A_MN, bool B_MN, bool PAIR = false, bool SPLIT = false, bool GRU = false>
__global__ void __launch_bounds__(GRU ? GRU_THREADS : THREADS, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                 const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmW2, const Params p) {
  static_assert(!PAIR || (!A_MN && !B_MN), "CTA pairs are built for the K-major forward form");
  static_assert(!SPLIT || (!A_MN && !B_MN), "the split form is built for the K-major forward form");
  using C = Cfg<BN, PAIR, SPLIT>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;                  // SWIZZLE_128B needs 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + C::STAGES * C::STAGE_BYTES;       // barrier block (256 B)
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (C::STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * C::STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * C::STAGES + 4);
  volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 4));
  float* params_smem = reinterpret_cast<float*>(base_ptr + C::STAGES * C::STAGE_BYTES + 256);
  float* stage_out = params_smem + C::PARAM_FLOATS;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // PAIR: the pair walks over (row-block pair, n_blk) tiles; this CTA owns row block 2*mp + rank of each
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const bool lead = crank == 0;
  const int num_tiles = p.tile_end;
  const int tile0 = p.tile_begin + (PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x);
  const int tile_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_m = [&](int tile) { return PAIR ? 2 * (tile / p.tiles_n) + (int)crank : tile / p.tiles_n; };
  const int num_kb = (p.K + BK - 1) / BK;      // K tail: TMA zero-fills both operands

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    if constexpr (SPLIT) { tma_prefetch_desc(&tmA2); tma_prefetch_desc(&tmW2); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    // accumulator release: 128 epilogue threads, or (PAIR) one arrive per epilogue warp of both CTAs on the leader's barrier
    for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), GRU ? 2 * GRU_EPI_WARPS : (PAIR ? 8 : 128)); }
    fence_barrier_init();
  }
  if (warp == 2) { if constexpr (PAIR) tmem_alloc_2cta(tmem_slot, C::TMEM_COLS); else tmem_alloc(tmem_slot, C::TMEM_COLS); }
  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();              // PDL: the next kernel may set itself up behind us ...
  griddep_wait();                // ... and we touch global memory only after our predecessor is complete

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int gt = GRU ? p.gru_t : 0; gt < (GRU ? p.gru_t_end : 1); ++gt)
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        const int m_blk = tile_m(tile), n_blk = tile % p.tiles_n;
        if constexpr (GRU) {
          if (p.gru_counter != nullptr && gt > p.gru_t && !(p.gru_debug & 2)) {      // h_{t-1}[rows of this CTA, :] comes from tiles_n CTAs
            const int target = (gt - p.gru_t) * p.tiles_n;
            unsigned int polls = 0;
            unsigned long long t0 = 0;
            while (ld_acquire_gpu(p.gru_counter + m_blk) < target) {
              if ((++polls & 0xFFFu) == 0) {             // a counter that never arrives must not hang the device: trap after 2 s
                unsigned long long now;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (t0 == 0) t0 = now;
                else if (now - t0 > 2000000000ull) __trap();
              }
            }
            fence_proxy_async_all();
          }
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::PLANES * C::A_BYTES;
          if constexpr (PAIR) {                      // both CTAs fill their own slots; bytes are counted on the leader's barrier
            if constexpr (GRU) {
              if (p.gru_debug & 8) {                 // timing experiment: no loads
                if (lead) mbar_arrive_expect_tx(full_bar(stage), 0u);
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                continue;
              }
            }
            if (lead) mbar_arrive_expect_tx(full_bar(stage), 2 * C::STAGE_BYTES);
            if constexpr (GRU) {                     // tmA: [2 buffers x 2 planes][B][H]
              const int z = ((gt - 1) & 1) * 2;
              tma_load_3d_2cta(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM, z);
              tma_load_3d_2cta(sa + C::A_BYTES, &tmA, full_bar(stage), kb * BK, m_blk * BM, z + 1);
            } else {
              tma_load_2d_2cta(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM);
              if constexpr (SPLIT) tma_load_2d_2cta(sa + C::A_BYTES, &tmA2, full_bar(stage), kb * BK, m_blk * BM);
            }
            tma_load_2d_2cta(sb, &tmW, full_bar(stage), kb * BK, n_blk * BN + (int)crank * (BN / 2));
            if constexpr (SPLIT)
              tma_load_2d_2cta(sb + C::B_BYTES, &tmW2, full_bar(stage), kb * BK, n_blk * BN + (int)crank * (BN / 2));
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), C::STAGE_BYTES);
          if constexpr (SPLIT) {
            tma_load_2d(sa + C::A_BYTES, &tmA2, full_bar(stage), kb * BK, m_blk * BM);
            tma_load_2d(sb + C::B_BYTES, &tmW2, full_bar(stage), kb * BK, n_blk * BN);
          }
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmA, full_bar(stage), kb * BK, m_blk * BM);
          } else {                                   // tensor [K rows, M cols]: 64-wide atoms of 64 k-rows
#pragma unroll
            for (int h = 0; h < BM / 64; ++h)
              tma_load_2d(sa + h * (BK * 128), &tmA, full_bar(stage), m_blk * BM + h * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmW, full_bar(stage), kb * BK, n_blk * BN);
          } else {
#pragma unroll
            for (int h = 0; h < BN / 64; ++h)
              tma_load_2d(sb + h * (BK * 128), &tmW, full_bar(stage), n_blk * BN + h * 64, kb * BK);
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0 && lead) {
      constexpr uint32_t idesc = (SPLIT ? make_idesc_f16(PAIR ? 2 * BM : BM, BN) : make_idesc_bf16(PAIR ? 2 * BM : BM, BN)) |
                                 (A_MN ? IDESC_A_MN_MAJOR : 0u) | (B_MN ? IDESC_B_MN_MAJOR : 0u);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int gt = GRU ? p.gru_t : 0; gt < (GRU ? p.gru_t_end : 1); ++gt)
      for (int tile = tile0; tile < num_tiles; tile += tile_step) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * C::ACC_STRIDE;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t sa = base + stage * C::STAGE_BYTES, sb = sa + C::PLANES * C::A_BYTES;
          if constexpr (SPLIT) {
            // main += A_hi·W_hi ; correction += A_hi·W_lo' + A_lo'·W_hi   (columns [BN, 2BN) of the tile)
            const uint64_t ah = make_sw128_kmajor_desc(sa), al = make_sw128_kmajor_desc(sa + C::A_BYTES);
            const uint64_t wh = make_sw128_kmajor_desc(sb), wl = make_sw128_kmajor_desc(sb + C::B_BYTES);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint32_t first = (kb | k) != 0;
              if constexpr (GRU) { if (p.gru_debug & 1) continue; }
              if constexpr (PAIR) {
                umma_bf16_2cta(d_tmem, ah + 2 * k, wh + 2 * k, idesc, first);
                umma_bf16_2cta(d_tmem + BN, ah + 2 * k, wl + 2 * k, idesc, first);
                umma_bf16_2cta(d_tmem + BN, al + 2 * k, wh + 2 * k, idesc, 1u);
              } else {
                umma_bf16(d_tmem, ah + 2 * k, wh + 2 * k, idesc, first);
                umma_bf16(d_tmem + BN, ah + 2 * k, wl + 2 * k, idesc, first);
                umma_bf16(d_tmem + BN, al + 2 * k, wh + 2 * k, idesc, 1u);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < (SPLIT ? 0 : BK / UMMA_K); ++k) {
            // K-major: advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom (+2 in the addr>>4 field);
            // MN-major: advance 16 k-rows = 2048 bytes, atoms BK*128 bytes apart
            const uint64_t adesc = A_MN ? make_sw128_mnmajor_desc(sa + k * 2048, BK * 128, 1024)
                                        : make_sw128_kmajor_desc(sa) + (uint64_t)(2 * k);
            const uint64_t bdesc = B_MN ? make_sw128_mnmajor_desc(sb + k * 2048, BK * 128, 1024)
                                        : make_sw128_kmajor_desc(sb) + (uint64_t)(2 * k);
            if constexpr (PAIR) umma_bf16_2cta(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
            else umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0);
          }
          if constexpr (PAIR) {                              // both CTAs' slots / accumulators
            umma_commit_2cta(empty_bar(stage), 0b11);
            if (kb == num_kb - 1) umma_commit_2cta(tfull_bar(acc), 0b11);
          } else {
            umma_commit(empty_bar(stage));                     // smem slot reusable when these MMAs retire
            if (kb == num_kb - 1) umma_commit(tfull_bar(acc)); // accumulator complete
          }
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == C::NACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= EPI_WARP0) {
    // ===== epilogue =====
    constexpr int EPI_THREADS_K = GRU ? GRU_EPI_WARPS * 32 : 128;
    const int q = (warp - EPI_WARP0) & 3;            // == warp % 4: TMEM lane quarter
    const int et = threadIdx.x - EPI_WARP0 * 32;     // 0..127 (GRU: 0..255)
    int acc = 0; uint32_t acc_phase = 0;
    int pbuf = 0;
    if constexpr (GRU) {
      // ===== GRU step epilogue (modules.py:153, torch GRU semantics): r = σ(gi_r + gh_r), z = σ(gi_z + gh_z),
      //       n = tanh(gi_n + r·gh_n), h' = (1 - z)·n + z·h; gh = the tile's accumulators + b_hh, gi = the token's table row.
      // Thread = (row, UT units of the tile).  One launch for all steps: the thread owns the same rows / units in every step
      // (at most two tiles per pair and step), so the fp32 state stays in registers between steps.
      static_assert(SPLIT && BN == 96 && PAIR, "GRU step: 32 units x [r|z|n] per pair tile, fp32-class mode");
      constexpr int U = BN / 3, UT = U / (GRU_EPI_WARPS / 4);        // 16 units per thread
      static_assert(UT == 16, "one 16-column TMEM load per gate");
      const int H = p.gru_H;
      const int c = ((warp - EPI_WARP0) >> 2) * UT;                   // this thread's first unit within the tile
      auto process_tile = [&](int gt, int tile, float (&h)[UT], bool load_h, bool store_h) {
        const int m_blk = tile_m(tile), n_blk = tile % p.tiles_n;
        const int u0 = n_blk * U;
        const int row = m_blk * BM + q * 32 + lane;
        const bool row_ok = row < p.M;
        float* ps = params_smem + pbuf * 3 * BN;
        for (int i = et; i < BN; i += EPI_THREADS_K) ps[BN + i] = p.gru_bhh[(i / U) * H + u0 + (i % U)];   // column = gate·U + unit
        long long tok = row_ok ? p.gru_tokens[(size_t)row * p.gru_T + gt] : 0;
        tok = tok < 0 ? 0 : (tok >= p.gru_rows ? p.gru_rows - 1 : tok);
        const float* gi = p.gru_table + (size_t)tok * 3 * H + u0 + c;
        if (row_ok) {
          // the table row does not depend on the GEMM: pull this thread's entries into L2 while the main loop runs
#pragma unroll
          for (int gate = 0; gate < 3; ++gate)
#pragma unroll
            for (int sct = 0; sct < UT / 8; ++sct) asm volatile("prefetch.global.L2 [%0];" ::"l"(gi + gate * H + sct * 8));
          if (load_h) {
            const float* hp = p.gru_h + (size_t)row * H + u0 + c;
#pragma unroll
            for (int j4 = 0; j4 < UT / 4; ++j4) {
              const float4 v = *(reinterpret_cast<const float4*>(hp) + j4);
              h[4 * j4] = v.x; h[4 * j4 + 1] = v.y; h[4 * j4 + 2] = v.z; h[4 * j4 + 3] = v.w;
            }
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS_K) : "memory");
        mbar_wait(tfull_bar(acc), acc_phase);
        tcgen05_fence_after();
        const uint32_t t_row = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
        const bool live = row_ok && !(p.gru_debug & 4);
        float4 g4[2][UT / 4];                                       // gi_r, gi_z: in flight together with the TMEM loads
        if (live) {
#pragma unroll
          for (int gate = 0; gate < 2; ++gate)
#pragma unroll
            for (int j4 = 0; j4 < UT / 4; ++j4) g4[gate][j4] = __ldg(reinterpret_cast<const float4*>(gi + gate * H) + j4);
        }
        float ghr[UT], ghz[UT], ghn[UT];
        {
          uint32_t mr[16], cr[16], mz[16], cz[16], mn[16], cn[16];
          tmem_ld_32x16(t_row + c, mr);          tmem_ld_32x16(t_row + BN + c, cr);
          tmem_ld_32x16(t_row + U + c, mz);      tmem_ld_32x16(t_row + BN + U + c, cz);
          tmem_ld_32x16(t_row + 2 * U + c, mn);  tmem_ld_32x16(t_row + BN + 2 * U + c, cn);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < UT; ++j) {
            ghr[j] = fmaf(__uint_as_float(cr[j]), 0x1p-11f, __uint_as_float(mr[j])) + ps[BN + c + j];
            ghz[j] = fmaf(__uint_as_float(cz[j]), 0x1p-11f, __uint_as_float(mz[j])) + ps[BN + U + c + j];
            ghn[j] = fmaf(__uint_as_float(cn[j]), 0x1p-11f, __uint_as_float(mn[j])) + ps[BN + 2 * U + c + j];
          }
        }
        tcgen05_fence_before();                                     // accumulator in registers: hand it back to the MMA thread
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
        if (live) {
          float4 gn4[UT / 4];                                       // gi_n: an L2 hit, in flight under the r / z arithmetic
#pragma unroll
          for (int j4 = 0; j4 < UT / 4; ++j4) gn4[j4] = __ldg(reinterpret_cast<const float4*>(gi + 2 * H) + j4);
          const float* gr = reinterpret_cast<const float*>(g4[0]);
          const float* gz = reinterpret_cast<const float*>(g4[1]);
          const float* gn = reinterpret_cast<const float*>(gn4);
#pragma unroll
          for (int j = 0; j < UT; ++j) {
            const float r = 1.f / (1.f + expf(-(gr[j] + ghr[j])));
            const float z = 1.f / (1.f + expf(-(gz[j] + ghz[j])));
            ghr[j] = r; ghz[j] = z;
          }
#pragma unroll
          for (int j = 0; j < UT; ++j) {
            const float n = tanhf(gn[j] + ghr[j] * ghn[j]);
            h[j] = (1.f - ghz[j]) * n + ghz[j] * h[j];
          }
          const bool final_step = gt == p.gru_T - 1;
          char* pl_dst = (final_step && p.gru_planes_last) ? reinterpret_cast<char*>(p.gru_planes_last)
                                                           : reinterpret_cast<char*>(p.gru_planes) + (size_t)(gt & 1) * 2 * p.out_plane;
          __half* oh = reinterpret_cast<__half*>(pl_dst) + (size_t)row * H + u0 + c;
          __half* ol = reinterpret_cast<__half*>(pl_dst + p.out_plane) + (size_t)row * H + u0 + c;
#pragma unroll
          for (int j8 = 0; j8 < UT / 8; ++j8) {
            uint4 rh, rl;
            uint32_t* ph = reinterpret_cast<uint32_t*>(&rh); uint32_t* pl = reinterpret_cast<uint32_t*>(&rl);
#pragma unroll
            for (int e = 0; e < 4; ++e) split_f16x2_pair(h[8 * j8 + 2 * e], h[8 * j8 + 2 * e + 1], ph[e], pl[e]);
            reinterpret_cast<uint4*>(oh)[j8] = rh;
            reinterpret_cast<uint4*>(ol)[j8] = rl;
          }
          if (store_h) {
            float* ho = ((final_step && p.gru_hout_last) ? p.gru_hout_last : p.gru_h) + (size_t)row * H + u0 + c;
#pragma unroll
            for (int j4 = 0; j4 < UT / 4; ++j4)
              reinterpret_cast<float4*>(ho)[j4] = make_float4(h[4 * j4], h[4 * j4 + 1], h[4 * j4 + 2], h[4 * j4 + 3]);
          }
        }
        if (p.gru_counter != nullptr && gt + 1 < p.gru_t_end) {
          // publish this tile's part of h_t (see gru_pair.cu: the release of the one thread that bumps the counter is
          // cumulative over the stores it observed through the barrier; TMA reads the planes, hence the proxy fence)
          asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS_K) : "memory");
          if (et == 0) {
            __threadfence();
            fence_proxy_async_all();
            red_release_gpu_add(p.gru_counter + m_blk, 1);
          }
        }
        if (++acc == C::NACC) { acc = 0; acc_phase ^= 1; }
        pbuf ^= 1;
      };
      if (p.gru_counter != nullptr) {
        // all steps in this launch, at most two tiles per pair and step: state in registers
        float h0[UT], h1[UT];
#pragma unroll
        for (int j = 0; j < UT; ++j) h0[j] = h1[j] = 0.f;
        for (int gt = p.gru_t; gt < p.gru_t_end; ++gt) {
          const bool first = gt == p.gru_t, last = gt + 1 == p.gru_t_end;
          if (tile0 < num_tiles) process_tile(gt, tile0, h0, first, last);
          if (tile0 + tile_step < num_tiles) process_tile(gt, tile0 + tile_step, h1, first, last);
        }
      } else {
        // one step per launch (any number of tiles per pair): state through global memory
        for (int gt = p.gru_t; gt < p.gru_t_end; ++gt)
          for (int tile = tile0; tile < num_tiles; tile += tile_step) {
            float h[UT];
#pragma unroll
            for (int j = 0; j < UT; ++j) h[j] = 0.f;
            process_tile(gt, tile, h, true, true);
          }
      }
    }
    for (int gt = GRU ? p.gru_t : 0; gt < (GRU ? 0 : 1); ++gt)
    for (int tile = tile0; tile < num_tiles; tile += tile_step, pbuf ^= 1) {
      const int m_blk = tile_m(tile), n_blk = tile % p.tiles_n;
      const int n0 = n_blk * BN;
      // stage this tile's per-column parameters; double buffered by tile parity (not by accumulator: a single-buffered
      // accumulator would let a fast warp overwrite what a slow one still reads; the barrier below orders buffer reuse)
      float* ps = params_smem + pbuf * 3 * BN;
      for (int c = et; c < BN; c += EPI_THREADS_K) {
        const int n = n0 + c;
        const bool ok = n < p.N;
        ps[c] = (ok && p.scale) ? p.scale[n] : 1.f;
        ps[BN + c] = (ok && p.bias) ? p.bias[n] : 0.f;
        ps[2 * BN + c] = (ok && p.logit_w) ? p.logit_w[n] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS_K) : "memory");
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();

      const int row = m_blk * BM + q * 32 + lane;
      const bool row_ok = row < p.M;
      const float* mul_row = p.mul ? p.mul + (size_t)((row_ok ? row : 0) / p.mul_row_div) * p.ld_mul : nullptr;
      const bool mul_vec = p.mul && ((p.ld_mul & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.mul) & 15) == 0);
      const float* add_row = p.add ? p.add + (size_t)((row_ok ? row : 0) / p.add_row_div) * p.ld_add : nullptr;
      const bool add_vec = p.add && ((p.ld_add & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.add) & 15) == 0);
      float part = 0.f;
      float best = 0.f; int bidx = -1;                 // fused argmax: first maximum of this row within the tile
      const uint32_t t_row = tmem_base + acc * C::ACC_STRIDE + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(t_row + c0, v);
        if constexpr (SPLIT) {
          uint32_t vc[32];
          tmem_ld_32x32(t_row + BN + c0, vc);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(fmaf(__uint_as_float(vc[j]), 0x1p-11f, __uint_as_float(v[j])));
        } else {
          tmem_ld_wait();
        }
        const int nbase = n0 + c0;
        if (nbase >= p.N) continue;
        float y[32];
        const bool full = nbase + 32 <= p.N;
#pragma unroll
        for (int j = 0; j < 32; ++j) y[j] = fmaf(__uint_as_float(v[j]), ps[c0 + j], ps[BN + c0 + j]);
        if (add_row && !p.add_after_act) {
          if (add_vec && full) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(add_row + nbase) + j4);
              y[4 * j4] += a4.x; y[4 * j4 + 1] += a4.y; y[4 * j4 + 2] += a4.z; y[4 * j4 + 3] += a4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nbase + j < p.N) y[j] += __ldg(add_row + nbase + j);
          }
        }
        if (p.relu) {
          if (p.leaky_slope == 0.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = fmaxf(y[j], 0.f);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) y[j] = y[j] > 0.f ? y[j] : p.leaky_slope * y[j];
          }
        }
        if (add_row && p.add_after_act) {
          if (add_vec && full) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 a4 = __ldg(reinterpret_cast<const float4*>(add_row + nbase) + j4);
              y[4 * j4] += a4.x; y[4 * j4 + 1] += a4.y; y[4 * j4 + 2] += a4.z; y[4 * j4 + 3] += a4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nbase + j < p.N) y[j] += __ldg(add_row + nbase + j);
          }
        }

        if (mul_row) {
          if (mul_vec && full) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const float4 m4 = __ldg(reinterpret_cast<const float4*>(mul_row + nbase) + j4);
              y[4 * j4] *= m4.x; y[4 * j4 + 1] *= m4.y; y[4 * j4 + 2] *= m4.z; y[4 * j4 + 3] *= m4.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nbase + j < p.N) y[j] *= __ldg(mul_row + nbase + j);
          }
        }
        if (p.mask && row_ok) {                                    // backward of a ReLU: pass where the saved output > 0
          if (p.mask_bf16) {
            const __nv_bfloat16* mk = reinterpret_cast<const __nv_bfloat16*>(p.mask) + (size_t)row * p.ld_mask + nbase;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nbase + j < p.N && !(__bfloat162float(mk[j]) > 0.f)) y[j] = 0.f;
          } else {
            const float* mk = reinterpret_cast<const float*>(p.mask) + (size_t)row * p.ld_mask + nbase;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (nbase + j < p.N && !(__ldg(mk + j) > 0.f)) y[j] = 0.f;
          }
        }
        if (p.sigmoid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) y[j] = __fdividef(1.f, 1.f + __expf(-y[j]));
        }
        if (p.amax_keys) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (nbase + j < p.N && (bidx < 0 || y[j] > best)) { best = y[j]; bidx = nbase + j; }
        }
        if (p.logit_w) {
#pragma unroll
          for (int j = 0; j < 32; ++j) part = fmaf(y[j], ps[2 * BN + c0 + j], part);   // logit_w = 0 beyond N
        } else if (row_ok) {
          if (p.out_split) {
            // fp16 plane pair for the next layer's A operand
            __half* oh = reinterpret_cast<__half*>(p.out) + (size_t)row * p.ldo + nbase;
            __half* ol = reinterpret_cast<__half*>(reinterpret_cast<char*>(p.out) + p.out_plane) + (size_t)row * p.ldo + nbase;
            if (full && ((p.ldo & 7) == 0) && (((reinterpret_cast<uintptr_t>(p.out) | p.out_plane) & 15) == 0)) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint4 rh, rl;
                uint32_t* ph = reinterpret_cast<uint32_t*>(&rh); uint32_t* pl = reinterpret_cast<uint32_t*>(&rl);
#pragma unroll
                for (int e = 0; e < 4; ++e) split_f16x2_pair(y[8 * j8 + 2 * e], y[8 * j8 + 2 * e + 1], ph[e], pl[e]);
                reinterpret_cast<uint4*>(oh)[j8] = rh;
                reinterpret_cast<uint4*>(ol)[j8] = rl;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (nbase + j < p.N) {
                  const __half h = __float2half_rn(y[j]);
                  oh[j] = h; ol[j] = __float2half_rn((y[j] - __half2float(h)) * 2048.f);
                }
            }
          } else if (p.out_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + nbase;
            if (full && ((p.ldo & 7) == 0) && ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0)) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint4 r;
                r.x = pack_bf16x2(y[8 * j8], y[8 * j8 + 1]); r.y = pack_bf16x2(y[8 * j8 + 2], y[8 * j8 + 3]);
                r.z = pack_bf16x2(y[8 * j8 + 4], y[8 * j8 + 5]); r.w = pack_bf16x2(y[8 * j8 + 6], y[8 * j8 + 7]);
                reinterpret_cast<uint4*>(o)[j8] = r;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) if (nbase + j < p.N) o[j] = __float2bfloat16_rn(y[j]);
            }
          }
        }
        if (!p.logit_w && !p.out_bf16 && !p.out_split) {
          // f32 output: thread = row would scatter 4-byte stores over 32 rows per instruction; transpose the
          // 32x32 chunk through shared memory so that every store instruction writes one 128-byte row segment
          float* st = stage_out + q * (32 * 33);
#pragma unroll
          for (int j = 0; j < 32; ++j) st[lane * 33 + j] = y[j];
          __syncwarp();
          const int row0 = m_blk * BM + q * 32;
          const int n = nbase + lane;
          float* o = reinterpret_cast<float*>(p.out) + (size_t)row0 * p.ldo + n;
          if (n < p.N) {
#pragma unroll 8
            for (int r = 0; r < 32; ++r)
              if (row0 + r < p.M) o[(size_t)r * p.ldo] = st[r * 33 + lane];
          }
          __syncwarp();
        }
      }
      if (p.logit_w && row_ok) reinterpret_cast<float*>(p.out)[(size_t)row * p.n_parts + n_blk] = part;
      {
        if (p.amax_keys) {
          // Row maxima of the tiles of one row block meet in a 64-bit atomicMax: the high word orders like the float
          // (+0 = -0), the low word is ~n, so among equal values the LOWEST column wins (torch.max's rule).  The tile
          // that arrives last at the row block's counter turns the keys into int64 labels.
          if (row_ok && bidx >= 0) {
            uint32_t b = __float_as_uint(best + 0.f);
            b ^= (b >> 31) ? 0xFFFFFFFFu : 0x80000000u;
            atomicMax(p.amax_keys + row, ((unsigned long long)b << 32) | (0xFFFFFFFFu - (uint32_t)bidx));
          }
          __threadfence();
          asm volatile("bar.sync 1, 128;" ::: "memory");
          volatile int* last_flag = reinterpret_cast<volatile int*>(base_ptr + C::STAGES * C::STAGE_BYTES + 8 * (2 * C::STAGES + 5));
          if (et == 0) *last_flag = (atomicAdd(p.amax_cnt + m_blk, 1) == p.tiles_n - 1);
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (*last_flag) {
            __threadfence();
            if (row_ok) {
              const unsigned long long key = *reinterpret_cast<volatile unsigned long long*>(p.amax_keys + row);
              p.amax_label[row] = (long long)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
            }
          }
        }
      }
      // release the accumulator stage
      tcgen05_fence_before();
      if constexpr (PAIR) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0);
      } else {
        mbar_arrive(tempty_bar(acc));
      }
      if (p.progress) {
        // this CTA's 128 x BN block of the output is written: publish it (all 128 epilogue threads stored rows of it)
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) red_release_gpu_add(p.progress + m_blk, 1);
      }
      if (++acc == C::NACC) { acc = 0; acc_phase ^= 1; }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();              // the peer may still read our slots / signal our barriers
  tcgen05_fence_after();
  if (warp == 2) { if constexpr (PAIR) tmem_dealloc_2cta(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS); }
}

// ---- host side -------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

int make_tensor_map_bf16(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld,
                         int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
                rows, cols, ld);
  return VQA_OK;
}

int make_tensor_map_bf16_nd(CUtensorMap* map, const void* ptr, int rank, const long long* dims,
                            const long long* strides_bytes, const int* box) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  if (rank < 2 || rank > 4 || box[0] != BK) return fail(VQA_ERR_INVALID, "make_tensor_map_bf16_nd: rank=%d box0=%d", rank, box[0]);
  cuuint64_t d[4]; cuuint64_t st[3]; cuuint32_t bx[4]; cuuint32_t es[4];
  for (int i = 0; i < rank; ++i) { d[i] = (cuuint64_t)dims[i]; bx[i] = (cuuint32_t)box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = (cuuint64_t)strides_bytes[i];
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VQA_ERR_CUDA, "cuTensorMapEncodeTiled (rank %d) failed with CUresult %d", rank, (int)r);
  return VQA_OK;
}

// MN-major operand: the tensor is [K rows, MN cols] row-major; box = {64 mn, 64 k-rows}
static int make_tensor_map_mn(CUtensorMap* map, const void* ptr, long long k_rows, long long mn_cols, long long ld) {
  return make_tensor_map_bf16(map, ptr, k_rows, mn_cols, ld, BK);
}

// [tile_begin, tile_end) of the call clamped to the kernel's tile count
static void tile_range(const vqa_linear_args& a, int total, int* begin, int* end) {
  int b = a.tile_begin > 0 ? a.tile_begin : 0;
  int e = (a.tile_end > 0 && a.tile_end < total) ? a.tile_end : total;
  if (b > e) b = e;
  *begin = b; *end = e;
}

// lo' planes of a split (VQA_F16X2) operand / output: explicit byte offset, or right behind the hi plane of `rows` rows
static size_t plane_offset(size_t given, long long rows, long long ld) { return given ? given : (size_t)rows * ld * 2; }

template <int BN, bool A_MN, bool B_MN, bool SPLIT = false>
static int launch(const vqa_linear_args& a, cudaStream_t s) {
  using C = Cfg<BN, false, SPLIT>;
  CUtensorMap tmA, tmW, tmA2, tmW2;
  int rc;
  if (A_MN) { if ((rc = make_tensor_map_mn(&tmA, a.d_A, a.K, a.M, a.lda))) return rc; }
  else if ((rc = make_tensor_map_bf16(&tmA, a.d_A, a.M, a.K, a.lda, BM))) return rc;
  if (B_MN) { if ((rc = make_tensor_map_mn(&tmW, a.d_W, a.K, a.N, a.ldw))) return rc; }
  else if ((rc = make_tensor_map_bf16(&tmW, a.d_W, a.N, a.K, a.ldw, BN))) return rc;
  tmA2 = tmA; tmW2 = tmW;
  if (SPLIT) {                                         // 2-byte elements: the bf16 maps serve fp16 planes as well
    if ((rc = make_tensor_map_bf16(&tmA2, (const char*)a.d_A + plane_offset(a.a_plane, a.M, a.lda), a.M, a.K, a.lda, BM))) return rc;
    if ((rc = make_tensor_map_bf16(&tmW2, (const char*)a.d_W + plane_offset(a.w_plane, a.N, a.ldw), a.N, a.K, a.ldw, BN))) return rc;
  }
  Params p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.scale = a.d_scale; p.bias = a.d_bias; p.relu = a.relu;
  p.mul = a.d_mul; p.ld_mul = a.ld_mul; p.mul_row_div = a.mul_row_div > 0 ? a.mul_row_div : 1;
  p.add = a.d_add; p.ld_add = a.ld_add; p.add_row_div = a.add_row_div > 0 ? a.add_row_div : 1;
  p.mask = a.d_mask; p.ld_mask = a.ld_mask; p.mask_bf16 = (a.mask_dtype == VQA_BF16);
  p.logit_w = a.d_logit_w; p.out = a.d_out; p.ldo = a.ldo; p.out_bf16 = (a.out_dtype == VQA_BF16);
  p.out_split = (a.out_dtype == VQA_F16X2); p.out_plane = plane_offset(a.out_plane, a.M, a.ldo);
  p.tiles_m = (a.M + BM - 1) / BM; p.tiles_n = (a.N + BN - 1) / BN; p.n_parts = p.tiles_n;
  p.leaky_slope = a.leaky_slope; p.add_after_act = a.add_after_act; p.sigmoid = a.sigmoid;
  p.amax_keys = nullptr; p.amax_cnt = nullptr; p.amax_label = nullptr;
  p.progress = a.d_progress;
  if (a.d_argmax_label) {
    VQA_REQUIRE(a.d_argmax_ws && !a.d_logit_w, "vqa_linear: fused argmax needs its zeroed workspace and the store form");
    p.amax_keys = (unsigned long long*)a.d_argmax_ws;
    p.amax_cnt = (int*)((char*)a.d_argmax_ws + align_up((size_t)a.M * 8, 256));
    p.amax_label = (long long*)a.d_argmax_label;
  }
  auto kern = linear_tc_kernel<BN, A_MN, B_MN, false, SPLIT>;
  static DeviceOnce attr;                            // per device, not per process
  const int dev = current_device();
  if (attr.need(dev)) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr.mark(dev);
  }
  tile_range(a, p.tiles_m * p.tiles_n, &p.tile_begin, &p.tile_end);
  const int tiles = p.tile_end - p.tile_begin;
  if (tiles <= 0) return VQA_OK;
  int grid = tiles < sm_count() ? tiles : sm_count();
  if (a.cta_limit > 0 && a.cta_limit < grid) grid = a.cta_limit;
  VQA_CUDA_CHECK(launch_pdl(kern, dim3(grid), dim3(THREADS), (size_t)C::SMEM_BYTES, s, tmA, tmW, tmA2, tmW2, p));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// 256 x 256 tiles on CTA pairs (tcgen05 cta_group::2) for the large K-major GEMMs.  VQA_ERR_UNSUPPORTED = the device cannot
// hold the pairs (or VQA_B200_GEMM_PAIR=0): the caller falls back to one CTA per tile.
// how many CTA pairs of the 256 x 256 kernel the current device holds at once (0 = pairs unusable)
template <bool SPLIT, int BN = 256>
static int pairs_resident_on_device() {
  using C = Cfg<BN, true, SPLIT>;
  auto kern = linear_tc_kernel<BN, false, false, true, SPLIT>;
  static DeviceInt cache;
  int& pairs_resident = cache.at(current_device());
  if (pairs_resident < 0) {
    const char* e = getenv("VQA_B200_GEMM_PAIR");
    pairs_resident = 0;
    if (!(e && e[0] == '0') &&
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (sm_count() / 2)); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) == cudaSuccess) pairs_resident = n;
    }
    (void)cudaGetLastError();
  }
  return pairs_resident;
}

template <bool SPLIT, int BN = 256>
static int launch_pair(const vqa_linear_args& a, cudaStream_t s) {
  using C = Cfg<BN, true, SPLIT>;
  auto kern = linear_tc_kernel<BN, false, false, true, SPLIT>;
  const int pairs_resident = pairs_resident_on_device<SPLIT, BN>();
  if (pairs_resident <= 0) return VQA_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmW, tmA2, tmW2;
  int rc;
  if ((rc = make_tensor_map_bf16(&tmA, a.d_A, a.M, a.K, a.lda, BM))) return rc;
  if ((rc = make_tensor_map_bf16(&tmW, a.d_W, a.N, a.K, a.ldw, BN / 2))) return rc;
  tmA2 = tmA; tmW2 = tmW;
  if (SPLIT) {
    if ((rc = make_tensor_map_bf16(&tmA2, (const char*)a.d_A + plane_offset(a.a_plane, a.M, a.lda), a.M, a.K, a.lda, BM))) return rc;
    if ((rc = make_tensor_map_bf16(&tmW2, (const char*)a.d_W + plane_offset(a.w_plane, a.N, a.ldw), a.N, a.K, a.ldw, BN / 2))) return rc;
  }
  Params p;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.scale = a.d_scale; p.bias = a.d_bias; p.relu = a.relu;
  p.mul = a.d_mul; p.ld_mul = a.ld_mul; p.mul_row_div = a.mul_row_div > 0 ? a.mul_row_div : 1;
  p.add = a.d_add; p.ld_add = a.ld_add; p.add_row_div = a.add_row_div > 0 ? a.add_row_div : 1;
  p.mask = a.d_mask; p.ld_mask = a.ld_mask; p.mask_bf16 = (a.mask_dtype == VQA_BF16);
  p.logit_w = a.d_logit_w; p.out = a.d_out; p.ldo = a.ldo; p.out_bf16 = (a.out_dtype == VQA_BF16);
  p.out_split = (a.out_dtype == VQA_F16X2); p.out_plane = plane_offset(a.out_plane, a.M, a.ldo);
  p.tiles_m = (a.M + BM - 1) / BM; p.tiles_n = (a.N + BN - 1) / BN; p.n_parts = p.tiles_n;
  p.leaky_slope = a.leaky_slope; p.add_after_act = a.add_after_act; p.sigmoid = a.sigmoid;
  p.amax_keys = nullptr; p.amax_cnt = nullptr; p.amax_label = nullptr;
  p.progress = a.d_progress;
  if (a.d_argmax_label) {
    VQA_REQUIRE(a.d_argmax_ws && !a.d_logit_w, "vqa_linear: fused argmax needs its zeroed workspace and the store form");
    p.amax_keys = (unsigned long long*)a.d_argmax_ws;
    p.amax_cnt = (int*)((char*)a.d_argmax_ws + align_up((size_t)a.M * 8, 256));
    p.amax_label = (long long*)a.d_argmax_label;
  }
  tile_range(a, ((p.tiles_m + 1) / 2) * p.tiles_n, &p.tile_begin, &p.tile_end);
  const int pair_tiles = p.tile_end - p.tile_begin;
  if (pair_tiles <= 0) return VQA_OK;
  int pairs = pair_tiles < pairs_resident ? pair_tiles : pairs_resident;
  if (a.cta_limit > 0 && a.cta_limit / 2 < pairs) pairs = a.cta_limit / 2 > 0 ? a.cta_limit / 2 : 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 2 : 1;
  VQA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmA2, tmW2, p));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

// Tile shape of a call.  Large GEMMs (a tile for every SM at BN = 256) are tensor-bound: widest tile, on CTA pairs when
// there are at least two 256-wide tiles per SM (only whole 256-wide N tiles — the shapes of the path: N = 1024, 6144;
// pairs on narrower tiles for the small-M layers were measured: no gain, those layers are bound by launch / fill / drain
// latency, not by operand ingest).  Small-M layers are bound by what one SM can pull in per tile — (BM + BN)·K operand
// bytes — times the number of waves, so pick the BN that minimises waves x (BM + BN); ties go to the wider tile.
// Measured on the 3129-way classifier (M = 1024, K = 2048): BN = 128 -> 200 tiles = 2 waves, 30 us; BN = 192 -> 136
// tiles = 1 wave.
struct Plan { int bn; bool pair; int tiles; };
static Plan plan_tiles(const vqa_linear_args& a, bool mn_major) {
  const int tiles_m = (a.M + BM - 1) / BM;
  const int sms = sm_count();
  Plan pl{256, false, 0};
  if (a.d_logit_w || tiles_m * ((a.N + 255) / 256) >= sms) {
    if (!mn_major && tiles_m * ((a.N + 255) / 256) >= 2 * sms && a.N % 256 == 0 && !a.d_argmax_label &&
        (a.dtype == VQA_F16X2 ? pairs_resident_on_device<true>() : pairs_resident_on_device<false>()) > 0)
      pl.pair = true;
  } else {
    int best_cost = 1 << 30;
    for (int bn : {256, 192, 128, 64}) {
      if (bn == 192 && mn_major) continue;           // MN-major operand tiles come in 64-wide atoms: 256 / 128 / 64 only
      const int tiles = tiles_m * ((a.N + bn - 1) / bn);
      const int cost = ((tiles + sms - 1) / sms) * (BM + bn);
      if (cost < best_cost) { best_cost = cost; pl.bn = bn; }
    }
  }
  pl.tiles = (pl.pair ? (tiles_m + 1) / 2 : tiles_m) * ((a.N + pl.bn - 1) / pl.bn);
  return pl;
}

// split (fp32-class) form: K-major only
static int launch_split(const vqa_linear_args& a, cudaStream_t s) {
  const Plan pl = plan_tiles(a, false);
  if (pl.pair) {
    const int rc = launch_pair<true>(a, s);
    if (rc != VQA_ERR_UNSUPPORTED) return rc;
  }
  // Small-M layers (one wave): a stage of the split form is two planes of both operands, so a single CTA holds only 2-4
  // stages and the MMAs wait for TMA (measured: the 3129-way classifier at 49 us against 19 us of MMA time).  On CTA pairs
  // a CTA stages half of the W tile — 3 to 5 stages — at the same number of busy SMs: the narrowest tile whose pair tiles
  // still fit the device in one wave.
  static int small_pairs = -1;
  if (small_pairs < 0) { const char* e = getenv("VQA_B200_SPLIT_SMALL_PAIR"); small_pairs = (e && e[0] == '0') ? 0 : 1; }
  const int tiles_m = (a.M + BM - 1) / BM;
  if (small_pairs && !pl.pair && tiles_m >= 2 && !a.d_progress && !a.d_logit_w /* logit parts are 256 wide */ && a.tile_begin == 0 && a.tile_end == 0 && a.cta_limit == 0) {
    const int mp = (tiles_m + 1) / 2;
    const int cap = pairs_resident_on_device<true, 64>();
    int rc = VQA_ERR_UNSUPPORTED;
    if (cap > 0) {
      if (mp * ((a.N + 63) / 64) <= cap) rc = launch_pair<true, 64>(a, s);
      else if (mp * ((a.N + 127) / 128) <= cap) rc = launch_pair<true, 128>(a, s);
      else if (mp * ((a.N + 191) / 192) <= cap) rc = launch_pair<true, 192>(a, s);
      else if (mp * ((a.N + 255) / 256) <= cap) rc = launch_pair<true, 256>(a, s);
    }
    if (rc != VQA_ERR_UNSUPPORTED) return rc;
  }
  switch (pl.bn) {
    case 256: return launch<256, false, false, true>(a, s);
    case 192: return launch<192, false, false, true>(a, s);
    case 128: return launch<128, false, false, true>(a, s);
    default: return launch<64, false, false, true>(a, s);
  }
}

template <bool A_MN, bool B_MN>
static int launch_bn(const vqa_linear_args& a, cudaStream_t s) {
  const Plan pl = plan_tiles(a, A_MN || B_MN);
  if constexpr (!A_MN && !B_MN) {
    if (pl.pair) {
      const int rc = launch_pair<false>(a, s);
      if (rc != VQA_ERR_UNSUPPORTED) return rc;
    }
  }
  switch (pl.bn) {
    case 256: return launch<256, A_MN, B_MN>(a, s);
    case 192:
      if constexpr (!A_MN && !B_MN) return launch<192, A_MN, B_MN>(a, s);
      else return launch<128, A_MN, B_MN>(a, s);
    case 128: return launch<128, A_MN, B_MN>(a, s);
    default: return launch<64, A_MN, B_MN>(a, s);
  }
}

// One step of the fp32-class GRU as ONE kernel: h_{t-1}·W_hhᵀ on CTA pairs (256 x 192 tiles of the gate-interleaved
// packed W_hh: 64 units x [r|z|n]) with the gate update as the epilogue.  VQA_ERR_UNSUPPORTED when the device cannot hold
// the pairs (the caller then runs the GEMM and the gate kernel separately).
static int gru_step_split(const GruStepSplit& g, cudaStream_t s) {
  // 32 units x [r|z|n] per tile: at B = 1024 every CTA pair owns TWO tiles of a step (two different row blocks), so the
  // tensor core runs one tile's MMAs while the 8 epilogue warps do the other tile's gate update (measured by switching
  // pieces off, 64-unit tiles: the gate update was 11 of a step's 28 us, the MMAs 7, and they ran one after the other)
  constexpr int BN = 96;
  using C = Cfg<BN, true, true>;
  auto kern = linear_tc_kernel<BN, false, false, true, true, true>;
  static DeviceInt cache;
  int& pairs_resident = cache.at(current_device());
  if (pairs_resident < 0) {
    pairs_resident = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * (sm_count() / 2)); cfg.blockDim = dim3(GRU_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, (const void*)kern, &cfg) == cudaSuccess) pairs_resident = n;
    }
    (void)cudaGetLastError();
  }
  if (pairs_resident <= 0 || g.H % 32 != 0) return VQA_ERR_UNSUPPORTED;
  CUtensorMap tmA, tmW, tmW2;
  int rc;
  const size_t a_plane = (size_t)g.B * g.H * 2, w_plane = (size_t)3 * g.H * g.H * 2;
  {
    const long long dims[3] = {g.H, g.B, 4}, strides[2] = {2LL * g.H, 2LL * g.H * g.B};
    const int box[3] = {BK, BM, 1};
    if ((rc = make_tensor_map_bf16_nd(&tmA, g.h_planes, 3, dims, strides, box))) return rc;
  }
  if ((rc = make_tensor_map_bf16(&tmW, g.wh_packed_planes, 3 * g.H, g.H, g.H, BN / 2))) return rc;
  if ((rc = make_tensor_map_bf16(&tmW2, (const char*)g.wh_packed_planes + w_plane, 3 * g.H, g.H, g.H, BN / 2))) return rc;
  Params p{};
  p.M = g.B; p.N = 3 * g.H; p.K = g.H;
  p.mul_row_div = 1; p.add_row_div = 1;
  p.ldo = g.H; p.out_split = 1; p.out_plane = a_plane;
  p.tiles_m = (g.B + BM - 1) / BM; p.tiles_n = 3 * g.H / BN; p.n_parts = p.tiles_n;
  p.gru_tokens = (const long long*)g.tokens; p.gru_table = g.gi_table; p.gru_bhh = g.b_hh; p.gru_h = g.h;
  p.gru_hout_last = g.h_out_last; p.gru_planes = g.h_planes; p.gru_planes_last = g.h_planes_last;
  p.gru_T = g.T; p.gru_H = g.H; p.gru_rows = g.ntoken_rows;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("VQA_B200_GRUS_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.gru_debug = dbg;
  p.tile_begin = 0; p.tile_end = ((p.tiles_m + 1) / 2) * p.tiles_n;
  // every tile on its own CTA pair, all of them resident: ONE launch for all steps, steps chained by the row-block
  // counters (VQA_B200_GRU_SPLIT_PERSIST=0: one launch per step).  Otherwise one launch per step.
  static int persist_env = -1;
  if (persist_env < 0) {                     // Nsight Compute cannot launch a cooperative cluster kernel: per-step launches there
    const char* e = getenv("VQA_B200_GRU_SPLIT_PERSIST");
    persist_env = e ? ((e[0] == '0') ? 0 : 1) : (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") ? 0 : 1);
  }
  // one launch for all steps when every CTA pair (all of them resident) owns at most two tiles of a step
  int pairs = p.tile_end <= pairs_resident ? p.tile_end : (p.tile_end + 1) / 2;
  const bool persist = persist_env && pairs <= pairs_resident && g.counter != nullptr && p.tiles_m <= 62 && g.t_end - g.t > 1;
  if (!persist) pairs = p.tile_end < pairs_resident ? p.tile_end : pairs_resident;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs); cfg.blockDim = dim3(GRU_THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES; cfg.stream = s;
  cudaLaunchAttribute at[2];
  at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  if (persist) {
    // one counter per 128-row block of every PAIR: with an odd number of row blocks the last pair's second CTA has no
    // rows of its own but still polls and bumps its counter
    VQA_CUDA_CHECK(cudaMemsetAsync(g.counter, 0, sizeof(int) * 2 * ((p.tiles_m + 1) / 2), s));
    p.gru_counter = g.counter; p.gru_t = g.t; p.gru_t_end = g.t_end;
    at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;      // the runtime checks co-residency
    cfg.numAttrs = 2;
    VQA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmA, tmW2, p));
    VQA_LAUNCH_CHECK();
    return VQA_OK;
  }
  at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  for (int t = g.t; t < g.t_end; ++t) {
    p.gru_counter = nullptr; p.gru_t = t; p.gru_t_end = t + 1;
    VQA_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmW, tmA, tmW2, p));
    VQA_LAUNCH_CHECK();
  }
  return VQA_OK;
}

}  // namespace tc

int linear_tc_gru_step(const GruStepSplit& g, cudaStream_t s) { return tc::gru_step_split(g, s); }

int linear_tc_part_width() { return 256; }

int linear_tc_tile_count(const vqa_linear_args& a) {
  if (a.M <= 0) return 0;
  return tc::plan_tiles(a, a.trans_a || a.trans_w).tiles;
}
int linear_tc_tiles_n(const vqa_linear_args& a) {
  if (a.M <= 0) return 0;
  const int bn = tc::plan_tiles(a, a.trans_a || a.trans_w).bn;
  return (a.N + bn - 1) / bn;
}

int linear_tc(const vqa_linear_args& a, cudaStream_t s) {
  VQA_REQUIRE(a.lda % 8 == 0 && a.ldw % 8 == 0 && (uintptr_t)a.d_A % 16 == 0 && (uintptr_t)a.d_W % 16 == 0,
              "vqa_linear(bf16): TMA needs 16-byte aligned rows (lda=%d ldw=%d)", a.lda, a.ldw);
  VQA_REQUIRE(!(a.trans_a && !a.trans_w), "vqa_linear(bf16): trans_a without trans_w is not built");
  if (a.M == 0) return VQA_OK;
  if (a.dtype == VQA_F16X2) {
    VQA_REQUIRE(!a.trans_a && !a.trans_w && !a.d_mask, "vqa_linear(f16x2): forward (K-major) form only");
    return tc::launch_split(a, s);
  }
  if (a.trans_a) return tc::launch_bn<true, true>(a, s);       // dW = dYᵀ·X
  if (a.trans_w) return tc::launch_bn<false, true>(a, s);      // dX = dY·W
  return tc::launch_bn<false, false>(a, s);
}

}  // namespace vqa
