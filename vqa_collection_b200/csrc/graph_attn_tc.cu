// k9+k10 on tcgen05 — relation-masked K×K graph attention for one CorrelatedGraphConv
// layer (bf16, sm_100a), consuming the merged wide projection Y = x·[W0+W1 ; W2 ; WbᵀWa]ᵀ.
//
// Reference: gcn.py:93-107 (DirectedGraphConv.conv incl. the label-bias gather),
// modules.py:86-95 (DotProduct), gcn.py:119-128 (relation_alpha: ReLU, adj·α, softmax over
// dim=1 = the ROW index), gcn.py:152-168 (forward), gcn.py:211-212 (ReLU after the layer),
// predictor.py:85 (Σ_K) when vsum is requested.
//
// Algebra (f_i = a_i·x_i with a = top-down attention; P=(W0+W1)x, S=W2x, Q=(WbᵀWa)x):
//   dot_ij = a_i a_j (Q_i·x_j) + a_i (x_i·Waᵀbb) + a_j (x_j·Wbᵀba) + ba·bb
//   α      = softmax_i( Σ_k adj_ik · ReLU(dot_kj) )
//   out_i  = ReLU( Σ_j α_ij conv_j ),  conv_j = a_j S_j + Σ_k adj_jk a_k P_k + Σ_l hist_jl bias_l
//          = ReLU( C_i · [P ; S ; bias] )      with the 36×96 coefficient matrix
//            C = [ (α·adj)·diag(a) | α·diag(a) | α·hist ]
// so the layer is two tensor-core contractions per image with a tiny K×K stage between them:
//   phase 1  D1[36,36] = Q·xᵀ (K = V) and D2[36,2] = x·[Waᵀbb, Wbᵀba]       tcgen05.mma, both K-major
//   phase 2  α and C in shared memory (fp32), C written as the bf16 K-major B operand
//   phase 3  outᵀ[128 ch, 36] = [P;S;bias]ᵀ·Cᵀ per 128-channel tile             tcgen05.mma, A is
//            M-major (channels contiguous): the TMA box {64 ch, rows} of Y IS the operand tile
// One persistent CTA per SM loops over images; warp 0 = TMA producer (one 24 KB slot ring for both
// phases), warp 1 = MMA issuer, warp 2 = TMEM allocator, warps 4..7 = phase 0/2 math (the K×K stage),
// warps 8..11 = phase-3 epilogue (thread = channel).  The image loop is software pipelined two deep:
// phase 1 of image n+1 (loads + MMAs into the second D1/D2 buffer) and phase 2 of image n+1 (into the
// second coefficient tile) run while the tensor core and the phase-3 warps work on image n, so neither
// the K×K stage nor the epilogue stalls the load stream (with one warp set doing phase 2 and phase 3 in
// turn the kernel ran at 0.68 of HBM with nothing saturated).
// Per image HBM traffic = Q, x, P, S (4 × 147 KB) + out: the kernel is HBM-bound by design.
#include <stdlib.h>

#include "tc_common.cuh"

namespace vqa {
namespace gat {

using namespace tc;

constexpr int GK = 36;                       // regions per image
constexpr int NPAD = 48;                     // regions padded to the MMA N granularity (16)
constexpr int P_ROWS = 40, S_ROWS = 40, LB_ROWS = 16;
constexpr int KROWS = P_ROWS + S_ROWS + LB_ROWS;          // 96 contraction rows of phase 3
constexpr int S_OFF = P_ROWS, LB_OFF = P_ROWS + S_ROWS;    // k index of the S / bias blocks
constexpr int CH = 128;                      // channels per phase-3 tile (MMA M)
constexpr int MAXL = 16;
constexpr int ATOM_BYTES = KROWS * 128;      // one 64-channel atom of the phase-3 A tile (12 KB)
constexpr int SLOT_BYTES = 2 * ATOM_BYTES;   // 24 KB
// phase-1 operands of one 64-channel k-block: Q 40 rows (5 KB) + x 40 rows (5 KB) + wvec 16 rows (2 KB) = 12 KB, TWO k-blocks
// per 24 KB ring slot.  (With one k-block per slot the ring carried 17 KB per slot on average and the kernel was bound by
// bytes in flight / HBM latency: 7 slots x 17 KB / 2.8 us = 43 GB/s per SM, 0.68 of HBM.)  The MMAs read 48 x-rows (N = 48):
// rows 40..47 are the first rows of the wvec tile — finite garbage that only reaches the unused columns 40..47 of D1.
constexpr int GROWS = 40;
constexpr int G_BYTES = 2 * GROWS * 128 + 16 * 128;       // 12 KB per k-block
static_assert(2 * G_BYTES == 2 * 12 * 1024, "two phase-1 k-blocks fill one ring slot");
// Slot layouts:
//   phase 1: two k-blocks, each [Q 40 rows | x 40 rows | wvec 16 rows] (5+5+2 KB) — ONE tcgen05.mma per k-step reads the
//            128 rows from Q on as A and the 64 rows from x on as B: D[0..39][0..39] = Q·xᵀ and D[40..79][40..55] =
//            x·[Waᵀbb, Wbᵀba]ᵀ land in one 64-column accumulator (two MMAs per k-step before; the kernel pays ≈ 50 ns per
//            MMA instruction whatever its N, measured by switching them off: 135.6 -> 122.1 us without the second one).
//            Rows 96..127 of A / 56..63 of B are whatever follows in shared memory — finite, and they only reach rows /
//            columns of D that nobody reads.
//   phase 3: [atom 0: P 40 rows, S 40 rows | atom 1: P, S | bias atom 0 | bias atom 1] (10+10+2+2 KB), one 4-D box for
//            the P and S rows of both 64-channel atoms, one 3-D box for the bias rows
constexpr int G_TILE = GROWS * 128;                       // one 40-row k-block tile (5 KB)
constexpr int G_Q_OFF = 0, G_X_OFF = G_TILE, G_W_OFF = 2 * G_TILE;
constexpr int U_ROW0 = 40, U_COL0 = 40;                   // D rows / columns of the x·wvec block
constexpr int PS_ATOM_BYTES = (P_ROWS + S_ROWS) * 128;    // 10 KB: the P and S rows of one 64-channel atom
constexpr int LB_SLOT_OFF = 2 * PS_ATOM_BYTES, LB_ATOM_BYTES = LB_ROWS * 128;
static_assert(LB_SLOT_OFF + 2 * LB_ATOM_BYTES == 2 * ATOM_BYTES, "phase-3 slot");
constexpr int STAGES = 7;
constexpr int BT_CHUNK = NPAD * 128;         // one 64-k chunk of the coefficient tile (6 KB)
constexpr int BT_BYTES = 2 * BT_CHUNK;
constexpr int BT_BUFS = 2;                   // coefficient tile double buffered over images
// warps 4..11: phase 0/2 (the K×K stage);  warps 12..15: phase-3 epilogue.  The K×K stage is the per-image critical
// path of a CTA (36 k warp-instructions of dependent shared-memory arithmetic per image): with 4 warps — one per
// scheduler, IPC 0.25 — it took ≈ 19 us per image whatever the memory system did (the kernel ran at the same 31-34 GB/s
// per SM from HBM on 148 SMs and from L2 on 16); 8 warps halve the work per thread and give every scheduler two warps.
constexpr int P2_WARPS = 8, P2_THREADS = P2_WARPS * 32;
constexpr int EPI_WARP0 = 4, EPI3_WARP0 = EPI_WARP0 + P2_WARPS, EPI_WARPS = P2_WARPS + 4, EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_WARP0 * 32 + EPI_THREADS;      // 512
constexpr int TMEM_COLS = 512;
constexpr int COL_G = 0, G_STRIDE = 64, COL_OUT = 128, OUT_STRIDE = 64, OUT_BUFS = 4;

constexpr int GP = GK + 4;                   // row pitch of A0 / Al: 16-byte aligned rows for float4 access
struct P2 {
  float G[GK][GK + 1];                       // Q·xᵀ
  alignas(16) float A0[GK][GP];              // ReLU(dot)
  alignas(16) float Al[GK][GP];              // adj·A0 → α
  float hist[GK][MAXL];
  float a[GK], ua[GK], ub[GK];
  unsigned long long adj[GK], adjT[GK];      // row masks (bit k: label_ik≠0), column masks (bit j: label_jk≠0)
};

constexpr int BAR_OFF = STAGES * SLOT_BYTES + BT_BUFS * BT_BYTES;
constexpr int P2_OFF = BAR_OFF + 256;
constexpr int SMEM_BYTES = 1024 + P2_OFF + (int)sizeof(P2);

struct Params {
  int B, V;
  int rev;                     // walk the images in descending order (the tail of Y is what the GEMM left in L2)
  const float* att; const uint8_t* labels; int num_labels; float c0;
  __nv_bfloat16* out; __nv_bfloat16* vsum; float* alpha;
  // chase mode: Y is being produced by a GEMM running beside this kernel; counter m of `progress` reaches
  // `progress_target` when rows [128m, 128m + 128) of Y are complete (gemm_tc.cu)
  const int* progress; int progress_target;
  int debug;                   // VQA_B200_GAT_DEBUG bits (timing experiments only, results are wrong):
                               // 2 = no phase-1 MMAs, 4 = no phase-3 MMAs, 8 = no K×K arithmetic
};

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(P2_THREADS) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
graph_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmX,
                          const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmPS,
                          const __grid_constant__ CUtensorMap tmLB, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bt = base + STAGES * SLOT_BYTES;               // coefficient tile (B operand of phase 3)
  uint8_t* bt_ptr = base_ptr + STAGES * SLOT_BYTES;
  const uint32_t bars = base + BAR_OFF;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto gfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + b); };
  auto cfull_bar = [&](int b) { return bars + 8u * (2 * STAGES + 2 + b); };          // coefficient tile b written
  auto cempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 4 + b); };         // ... and consumed by phase 3
  auto ofull_bar = [&](int b) { return bars + 8u * (2 * STAGES + 6 + b); };
  auto oempty_bar = [&](int b) { return bars + 8u * (2 * STAGES + 6 + OUT_BUFS + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 6 + 2 * OUT_BUFS);
  volatile uint32_t* tmem_slot_ptr =
      reinterpret_cast<volatile uint32_t*>(base_ptr + BAR_OFF + 8 * (2 * STAGES + 6 + 2 * OUT_BUFS));
  P2& sm = *reinterpret_cast<P2*>(base_ptr + P2_OFF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int V = p.V;
  const int kb_g = V / BK;                   // phase-1 k-blocks
  const int n_cc = V / CH;                   // phase-3 channel tiles

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmPS); tma_prefetch_desc(&tmLB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
    mbar_init(gfull_bar(0), 1); mbar_init(gfull_bar(1), 1);
    for (int b = 0; b < BT_BUFS; ++b) { mbar_init(cfull_bar(b), 1); mbar_init(cempty_bar(b), 1); }
    for (int b = 0; b < OUT_BUFS; ++b) { mbar_init(ofull_bar(b), 1); mbar_init(oempty_bar(b), 128); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  if (warp >= EPI_WARP0) {
    // coefficient tile: rows 36..47 and the k padding stay zero for the whole kernel
    for (int i = threadIdx.x - EPI_WARP0 * 32; i < BT_BUFS * BT_BYTES / 16; i += EPI_THREADS)
      reinterpret_cast<uint4*>(bt_ptr)[i] = make_uint4(0, 0, 0, 0);
    fence_proxy_async();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  griddep_launch();              // PDL (common.cuh): next kernel sets up behind us; global memory only after the wait
  griddep_wait();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      auto load_g = [&](int img) {                   // phase-1 operands of one image
        const int row0 = (p.rev ? p.B - 1 - img : img) * GK;
        if (p.progress) {
          // first touch of this image's rows of Y: wait until the producer GEMM has finished the (one or two) 128-row
          // blocks that hold them, then make its generic-proxy stores visible to the TMA loads (async proxy)
          // (the 40-row TMA boxes reach 4 rows into the next image: those rows meet zero coefficients in phase 3, but
          // they must be finished — finite — data, so their row block is waited for as well)
          const int m_last = (p.B * GK - 1) / BM;
          const int m_lo = row0 / BM, m_hi = min((row0 + GROWS - 1) / BM, m_last);
          unsigned long long t0 = 0;
          for (int m = m_lo; m <= m_hi; ++m) {
            unsigned int polls = 0;
            while (ld_acquire_gpu(p.progress + m) < p.progress_target) {
              if ((++polls & 0x3FFu) == 0) {           // a producer that never runs must not hang the device
                const unsigned long long now = global_timer_ns();
                if (t0 == 0) t0 = now;
                else if (now - t0 > 2000000000ull) __trap();
              }
              __nanosleep(64);
            }
          }
          fence_proxy_async_all();
        }
        for (int kb = 0; kb < kb_g; kb += 2) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t slot = base + stage * SLOT_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), 2 * G_BYTES);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t sk = slot + h * G_BYTES;
            tma_load_3d(sk + G_Q_OFF, &tmQ, full_bar(stage), 0, row0, (2 * V) / BK + kb + h);
            tma_load_3d(sk + G_X_OFF, &tmX, full_bar(stage), 0, row0, kb + h);
            tma_load_3d(sk + G_W_OFF, &tmW, full_bar(stage), 0, 0, kb + h);                  // [Waᵀbb, Wbᵀba]
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      };
      if ((int)blockIdx.x < p.B) load_g(blockIdx.x);
      for (int img = blockIdx.x; img < p.B; img += gridDim.x) {
        if (img + (int)gridDim.x < p.B) load_g(img + gridDim.x);
        const int row0 = (p.rev ? p.B - 1 - img : img) * GK;
        for (int cc = 0; cc < n_cc; ++cc) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t slot = base + stage * SLOT_BYTES;
          mbar_arrive_expect_tx(full_bar(stage), SLOT_BYTES);
          tma_load_4d(slot, &tmPS, full_bar(stage), 0, row0, 0, cc * (CH / 64));          // P and S rows of both atoms
          tma_load_3d(slot + LB_SLOT_OFF, &tmLB, full_bar(stage), 0, 0, cc * (CH / 64));  // bias rows of both atoms
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc_g = make_idesc_bf16(128, 64);
      constexpr uint32_t idesc_o = make_idesc_bf16(128, NPAD) | IDESC_A_MN_MAJOR;
      int stage = 0; uint32_t phase = 0;
      auto mma_g = [&](uint32_t gbuf) {
        // phase 1 into D1/D2 buffer gbuf (free: its previous image's phase 2a finished before that image's c_full)
        const uint32_t dg = tmem_base + gbuf * G_STRIDE;
        for (int kb = 0; kb < kb_g; kb += 2) {
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t slot = base + stage * SLOT_BYTES;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t sk = slot + h * G_BYTES;
            const uint64_t ad = make_sw128_kmajor_desc(sk + G_Q_OFF), bd = make_sw128_kmajor_desc(sk + G_X_OFF);
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k)
              if (!(p.debug & 2)) umma_bf16(dg + COL_G, ad + 2 * k, bd + 2 * k, idesc_g, (kb | h | k) != 0);
          }
          umma_commit(empty_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(gfull_bar(gbuf));
      };
      uint32_t it = 0;
      if ((int)blockIdx.x < p.B) mma_g(0);
      for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
        if (img + (int)gridDim.x < p.B) mma_g((it + 1) & 1u);
        // phase 3 needs the coefficient tile of this image
        const uint32_t cb = it & 1u;
        mbar_wait(cfull_bar(cb), (it >> 1) & 1u);
        tcgen05_fence_after();
        const uint32_t btb = bt + cb * BT_BYTES;
        for (int cc = 0; cc < n_cc; ++cc) {
          const uint32_t tile = it * (uint32_t)n_cc + (uint32_t)cc;      // running tile index of this CTA
          const int buf = tile & (OUT_BUFS - 1);
          const uint32_t use = tile / OUT_BUFS;
          mbar_wait(oempty_bar(buf), (use & 1u) ^ 1u);
          tcgen05_fence_after();
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t slot = base + stage * SLOT_BYTES;
          const uint32_t d = tmem_base + COL_OUT + buf * OUT_STRIDE;
#pragma unroll
          for (int ks = 0; ks < KROWS / UMMA_K; ++ks) {
            // k rows 0..79 = the P and S rows (atoms 10 KB apart), k rows 80..95 = the bias rows (atoms 2 KB apart)
            const uint64_t ad = ks < (P_ROWS + S_ROWS) / UMMA_K
                                    ? make_sw128_mnmajor_desc(slot + ks * 2048, PS_ATOM_BYTES, 1024)
                                    : make_sw128_mnmajor_desc(slot + LB_SLOT_OFF, LB_ATOM_BYTES, 1024);
            const uint64_t bd = make_sw128_kmajor_desc(btb + (ks >> 2) * BT_CHUNK) + (uint64_t)(2 * (ks & 3));
            if (!(p.debug & 4)) umma_bf16(d, ad, bd, idesc_o, ks != 0);
          }
          umma_commit(empty_bar(stage));
          umma_commit(ofull_bar(buf));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(cempty_bar(cb));                   // the phase-2 warps may rewrite this coefficient tile
      }
    }
  } else if (warp >= EPI_WARP0) {
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int hf = warp >= EPI3_WARP0 ? 1 : 0;       // 0: phase 0/2 warps, 1: phase-3 epilogue warps
    const int et = threadIdx.x - EPI_WARP0 * 32;     // index within the phase-2 group
    // ---- phase 0: attention scalars, row adjacency masks, label histogram of one image
    auto phase0 = [&](int img_it) {
      const int img = p.rev ? p.B - 1 - img_it : img_it;
      // 4 threads per region row, 9 labels each (whole warps take part in the shuffles): adjacency bits and a packed
      // 8-bit-per-label histogram, combined over the 4 neighbouring lanes
      if (et < 160) {
        const bool valid = et < 4 * GK;
        const int r = valid ? (et >> 2) : 0, part = et & 3;
        const uint8_t* lr = p.labels + ((size_t)img * GK + r) * GK + part * (GK / 4);
        unsigned long long m = 0ull, h0 = 0ull, h1 = 0ull;
#pragma unroll
        for (int k = 0; k < GK / 4; ++k) {
          const unsigned l = valid ? (unsigned)__ldg(lr + k) : 0u;
          if (l != 0) m |= 1ull << (part * (GK / 4) + k);
          if (l < 8) h0 += 1ull << (8 * l);
          else if (l < MAXL) h1 += 1ull << (8 * (l - 8));
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
          m |= __shfl_xor_sync(0xffffffffu, m, o);
          h0 += __shfl_xor_sync(0xffffffffu, h0, o);
          h1 += __shfl_xor_sync(0xffffffffu, h1, o);
        }
        if (valid && part == 0) {
          sm.a[r] = p.att ? __ldg(p.att + (size_t)img * GK + r) : 1.f;
          sm.adj[r] = m;
#pragma unroll
          for (int l = 0; l < MAXL; ++l) {
            const unsigned cnt = (unsigned)(((l < 8 ? h0 >> (8 * l) : h1 >> (8 * (l - 8)))) & 0xffull);
            sm.hist[r][l] = (l < p.num_labels) ? (float)cnt : 0.f;
          }
        }
      }
    };
    uint32_t it = 0;
    if (hf == 0) {
    // ================= phase 0/2 warps =================
    if ((int)blockIdx.x < p.B) phase0(blockIdx.x);
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
      // ---- phase 2a: D1/D2 rows 0..35 → shared memory
      const uint32_t gbuf = it & 1u;
      mbar_wait(gfull_bar(gbuf), (it >> 1) & 1u);
      tcgen05_fence_after();
      if (q < 3 && warp < EPI_WARP0 + 4) {          // D rows 0..35 = Q·xᵀ, rows 40..75 = x·wvecᵀ: one warp per lane quarter
        uint32_t v[32], w16[16], u8[8];
        const uint32_t t_row = tmem_base + gbuf * G_STRIDE + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(t_row + COL_G, v);
        tmem_ld_32x16(t_row + COL_G + 32, w16);
        tmem_ld_32x8(t_row + COL_G + U_COL0, u8);
        tmem_ld_wait();
        const int r = q * 32 + lane;
        if (r < GK) {
#pragma unroll
          for (int j = 0; j < 32; ++j) sm.G[r][j] = __uint_as_float(v[j]);
#pragma unroll
          for (int j = 0; j < GK - 32; ++j) sm.G[r][32 + j] = __uint_as_float(w16[j]);
        }
        if (r >= U_ROW0 && r < U_ROW0 + GK) {
          sm.ua[r - U_ROW0] = __uint_as_float(u8[0]);
          sm.ub[r - U_ROW0] = __uint_as_float(u8[1]);
        }
      }
      tcgen05_fence_before();
      epi_bar();
      if (p.debug & 8) {                                 // timing experiment: hand the (stale) coefficient tile over at once
        const uint32_t cbd = it & 1u;
        if (it >= 2) mbar_wait(cempty_bar(cbd), ((it >> 1) - 1u) & 1u);
        epi_bar();
        if (et == 0) mbar_arrive(cfull_bar(cbd));
        continue;
      }
      // ---- 2b: α0 = ReLU(dot); column masks adjT from the row masks
      for (int e = et; e < GK * GK; e += P2_THREADS) {
        const int i = e / GK, j = e - i * GK;
        const float ai = sm.a[i], aj = sm.a[j];
        const float dot = ai * aj * sm.G[i][j] + ai * sm.ua[i] + aj * sm.ub[j] + p.c0;
        sm.A0[i][j] = fmaxf(dot, 0.f);
      }
      if (et >= P2_THREADS - GK) {
        const int k = et - (P2_THREADS - GK);
        unsigned long long m = 0ull;
#pragma unroll
        for (int j = 0; j < GK; ++j) m |= ((sm.adj[j] >> k) & 1ull) << j;
        sm.adjT[k] = m;
      }
      epi_bar();
      // ---- 2c: α1 = adj·α0
      for (int e = et; e < GK * (GK / 4); e += P2_THREADS) {          // thread = (row i, 4 columns): one LDS.128 per k
        const int i = e / (GK / 4), j4 = (e - i * (GK / 4)) * 4;
        const unsigned long long m = sm.adj[i];
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < GK; ++k) {
          const float4 v = *reinterpret_cast<const float4*>(&sm.A0[k][j4]);
          const bool on = (m >> k) & 1ull;
          s.x += on ? v.x : 0.f; s.y += on ? v.y : 0.f; s.z += on ? v.z : 0.f; s.w += on ? v.w : 0.f;
        }
        *reinterpret_cast<float4*>(&sm.Al[i][j4]) = s;
      }
      epi_bar();
      // ---- 2d: softmax over the row index i for every column j (4 threads per column, 9 rows each)
      {
        const bool valid = et < 2 * GK;                  // whole warps run the shuffles; the tail lanes idle along
        const int j = valid ? (et >> 1) : 0, part = et & 1;
        float x[GK / 2];
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < GK / 2; ++r) { x[r] = valid ? sm.Al[part * (GK / 2) + r][j] : 0.f; mx = fmaxf(mx, x[r]); }
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < GK / 2; ++r) { x[r] = expf(x[r] - mx); sum += x[r]; }
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        const float inv = 1.f / sum;
        __syncwarp();                                    // all reads of column j done before it is overwritten
        if (valid) {
#pragma unroll
          for (int r = 0; r < GK / 2; ++r) sm.Al[part * (GK / 2) + r][j] = x[r] * inv;
        }
      }
      epi_bar();
      // ---- 2e: coefficient matrix C [36, 96] → bf16, K-major 128B-swizzled B operand
      //      k 0..35: (α·adj)_ik a_k   |  k 40..75: α_ij a_j  |  k 80..91: (α·hist)_il
      const uint32_t cb = it & 1u;
      if (it >= 2) mbar_wait(cempty_bar(cb), ((it >> 1) - 1u) & 1u);     // phase 3 of image it-2 has read this tile
      uint8_t* btw = bt_ptr + cb * BT_BYTES;
      auto put2 = [&](int i, int k0, float c0v, float c1v) {
        const int chunk = k0 >> 6, kc = k0 & 63;
        const uint32_t off = chunk * BT_CHUNK + i * 128 + ((((kc * 2) >> 4) ^ (i & 7)) << 4) + ((kc * 2) & 15);
        *reinterpret_cast<uint32_t*>(btw + off) = pack_bf16x2(c0v, c1v);
      };
      for (int e = et; e < GK * (GK / 2); e += P2_THREADS) {          // P block
        const int i = e / (GK / 2), kp = e - i * (GK / 2);
        const unsigned long long m0 = sm.adjT[2 * kp], m1 = sm.adjT[2 * kp + 1];
        float c0v = 0.f, c1v = 0.f;
#pragma unroll
        for (int j = 0; j < GK; ++j) {
          const float al = sm.Al[i][j];
          c0v += ((m0 >> j) & 1ull) ? al : 0.f;
          c1v += ((m1 >> j) & 1ull) ? al : 0.f;
        }
        put2(i, 2 * kp, c0v * sm.a[2 * kp], c1v * sm.a[2 * kp + 1]);
      }
      for (int e = et; e < GK * (GK / 2); e += P2_THREADS) {          // S block
        const int i = e / (GK / 2), kp = e - i * (GK / 2);
        put2(i, S_OFF + 2 * kp, sm.Al[i][2 * kp] * sm.a[2 * kp], sm.Al[i][2 * kp + 1] * sm.a[2 * kp + 1]);
      }
      for (int e = et; e < GK * (MAXL / 2); e += P2_THREADS) {        // label-bias block
        const int i = e / (MAXL / 2), lp = e - i * (MAXL / 2);
        float c0v = 0.f, c1v = 0.f;
#pragma unroll
        for (int j = 0; j < GK; ++j) {
          const float al = sm.Al[i][j];
          c0v = fmaf(al, sm.hist[j][2 * lp], c0v);
          c1v = fmaf(al, sm.hist[j][2 * lp + 1], c1v);
        }
        put2(i, LB_OFF + 2 * lp, c0v, c1v);
      }
      if (p.alpha != nullptr)
        for (int e = et; e < GK * GK; e += P2_THREADS)
          p.alpha[(size_t)(p.rev ? p.B - 1 - img : img) * GK * GK + e] = sm.Al[e / GK][e % GK];
      fence_proxy_async();
      epi_bar();
      if (et == 0) mbar_arrive(cfull_bar(cb));
      if (img + (int)gridDim.x < p.B) phase0(img + gridDim.x);       // consumed after the next 2a barrier
    }
    } else {
    // ================= phase-3 epilogue warps: thread = channel; ReLU, Σ_i, store =================
    for (int img = blockIdx.x; img < p.B; img += gridDim.x, ++it) {
      for (int cc = 0; cc < n_cc; ++cc) {
        const uint32_t tile = it * (uint32_t)n_cc + (uint32_t)cc;
        const int buf = tile & (OUT_BUFS - 1);
        const uint32_t use = tile / OUT_BUFS;
        mbar_wait(ofull_bar(buf), use & 1u);
        tcgen05_fence_after();
        uint32_t v[32], w16[16];
        const uint32_t t_row = tmem_base + COL_OUT + buf * OUT_STRIDE + ((uint32_t)(q * 32) << 16);
        tmem_ld_32x32(t_row, v);
        tmem_ld_32x16(t_row + 32, w16);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(oempty_bar(buf));
        const int c = cc * CH + q * 32 + lane;
        float vs = 0.f;
        const int oimg = p.rev ? p.B - 1 - img : img;
        __nv_bfloat16* o = p.out ? p.out + (size_t)oimg * GK * V + c : nullptr;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float s = fmaxf(__uint_as_float(v[i]), 0.f);
          vs += s;
          if (o) o[(size_t)i * V] = __float2bfloat16_rn(s);
        }
#pragma unroll
        for (int i = 32; i < GK; ++i) {
          const float s = fmaxf(__uint_as_float(w16[i - 32]), 0.f);
          vs += s;
          if (o) o[(size_t)i * V] = __float2bfloat16_rn(s);
        }
        if (p.vsum) p.vsum[(size_t)oimg * V + c] = __float2bfloat16_rn(vs);
      }
    }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

}  // namespace gat

int graph_attention_tc(const vqa_graph_attention_args& a, cudaStream_t s) {
  using namespace gat;
  VQA_REQUIRE(a.d_Y && a.d_x && a.d_wvec && a.d_label_bias_lp && a.d_labels, "graph_attention(layout 1): NULL input");
  if (a.K != GK) return fail(VQA_ERR_UNSUPPORTED, "graph_attention: K=%d (only K=%d is built)", a.K, GK);
  VQA_REQUIRE(a.dtype == VQA_BF16, "graph_attention(layout 1): bf16 only");
  VQA_REQUIRE(a.V % CH == 0, "graph_attention(layout 1): V=%d must be a multiple of %d", a.V, CH);
  VQA_REQUIRE(a.num_labels >= 1 && a.num_labels <= MAXL, "graph_attention: num_labels=%d", a.num_labels);
  VQA_REQUIRE(a.ldy >= 3 * a.V && a.ldy % 8 == 0 && a.ldx >= a.V && a.ldx % 8 == 0, "graph_attention(layout 1): ldy=%d ldx=%d",
              a.ldy, a.ldx);
  if (a.B == 0) return VQA_OK;
  CUtensorMap tmQ, tmX, tmW, tmPS, tmLB;
  int rc;
  const long long rows = (long long)a.B * GK;
  {
    // every operand is a row-major matrix viewed as [64-column atoms][rows][64 cols]: one box = several atoms
    const int bq[3] = {BK, GROWS, 1}, bw1[3] = {BK, 16, 1}, bw[3] = {BK, 16, 2}, bps[4] = {BK, P_ROWS, 2, 2};
    const long long dq[3] = {BK, rows, a.ldy / BK}, sq[2] = {2LL * a.ldy, 2LL * BK};
    const long long dx[3] = {BK, rows, a.V / BK}, sx[2] = {2LL * a.ldx, 2LL * BK};
    const long long dw[3] = {BK, 16, a.V / BK}, sw[2] = {2LL * a.V, 2LL * BK};
    // P / S: [atoms][map: P = 0, S = 1 (V columns further)][rows][64 cols]
    const long long dps[4] = {BK, rows, 2, a.V / BK}, sps[3] = {2LL * a.ldy, 2LL * a.V, 2LL * BK};
    VQA_REQUIRE(a.ldy % BK == 0 && a.V % BK == 0, "graph_attention(layout 1): ldy and V must be multiples of %d", BK);
    if ((rc = tc::make_tensor_map_bf16_nd(&tmQ, a.d_Y, 3, dq, sq, bq))) return rc;
    if ((rc = tc::make_tensor_map_bf16_nd(&tmX, a.d_x, 3, dx, sx, bq))) return rc;
    if ((rc = tc::make_tensor_map_bf16_nd(&tmW, a.d_wvec, 3, dw, sw, bw1))) return rc;
    if ((rc = tc::make_tensor_map_bf16_nd(&tmPS, a.d_Y, 4, dps, sps, bps))) return rc;
    if ((rc = tc::make_tensor_map_bf16_nd(&tmLB, a.d_label_bias_lp, 3, dw, sw, bw))) return rc;
  }
  Params p;
  p.B = a.B; p.V = a.V; p.att = a.d_att; p.labels = a.d_labels; p.num_labels = a.num_labels; p.c0 = a.c0;
  p.out = (__nv_bfloat16*)a.d_out; p.vsum = (__nv_bfloat16*)a.d_vsum; p.alpha = a.d_alpha;
  p.progress = a.d_progress; p.progress_target = a.progress_target;
  static int dbg = -1;
  if (dbg < 0) { const char* e = getenv("VQA_B200_GAT_DEBUG"); dbg = e ? atoi(e) : 0; }
  p.debug = dbg;
  VQA_REQUIRE(!a.d_progress || a.progress_target > 0, "graph_attention(layout 1): d_progress needs progress_target > 0");
  // measured: no gain here (1089 vs 1085 us per ReGAT step) — Y is 3.6x the L2, the tail that survives is small
  static int rev = -1;
  if (rev < 0) { const char* e = getenv("VQA_B200_GAT_REVERSE"); rev = (e && e[0] == '1') ? 1 : 0; }
  p.rev = a.d_progress ? 0 : rev;                    // chase mode walks the images in the order the GEMM finishes them
  static DeviceOnce attr;                            // per device, not per process
  if (attr.need(current_device())) {
    VQA_CUDA_CHECK(cudaFuncSetAttribute(graph_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr.mark(current_device());
  }
  int grid = a.B < sm_count() ? a.B : sm_count();
  if (a.cta_limit > 0 && a.cta_limit < grid) grid = a.cta_limit;
  VQA_CUDA_CHECK(launch_pdl(graph_attention_tc_kernel, dim3(grid), dim3(THREADS), (size_t)SMEM_BYTES, s, tmQ, tmX, tmW, tmPS,
                            tmLB, p));
  VQA_LAUNCH_CHECK();
  return VQA_OK;
}

}  // namespace vqa
