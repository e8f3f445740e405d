// stk_gemm.cu — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear of the path and its autograd backward (reference call sites:
// stonkgs_model.py:62-73,204-217; arithmetic in HF modeling_bert.py:158-160,179-181,287-298,
// 330-356,456-468,471-485).
//
// Design (persistent; one CTA per SM, CTA pairs on 256 x 256 tiles whenever M > 128):
//   warps 0-7   epilogue: tcgen05.ld of this CTA's 128 x 256 fp32 accumulator (warp w: TMEM lane quarter w % 4, column
//               half w / 4), fused epilogue math, swizzled staging tile, TMA store / reduce-add
//   warp 8      TMA producer: A / B tiles -> shared-memory ring of 128B-swizzled boxes (PAIR: 6 stages of 32 KB, each CTA
//               stages its 128 rows of A and 128 of the 256 rows of B; both CTAs' loads complete on the even CTA's
//               mbarrier)
//   warp 9      TMEM allocator + MMA issuer: the whole converged warp executes one statement per k-block (elect.sync
//               inside it, four UMMA 128x256x16 / cta_group::2 256x256x16 + a probe of the next slot + the slot-free
//               commit), see umma_kblock_warp
//   warps 10-11 (LayerNorm epilogues only) I/O warps: residual in, pre-LN sum / normalised rows out, by TMA
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers over TWO accumulator stages
// (2 x 256 columns = all 512 TMEM columns) so the epilogue of tile i overlaps the MMAs of tile i+1, and a static
// persistent tile scheduler (grid = the SMs stk_set_sm_reserve leaves to persistent kernels).
// Operand layouts: K-major or MN-major for A and B independently (UMMA descriptor major bits), so
// forward, dgrad and wgrad read the tensors where they lie — no transposed copies.
// Tails in M, N and K are handled by TMA (zero fill on load, clipping on store).
#include <atomic>
#include <stdlib.h>

#include "stk_common.cuh"
#include "stk_host.h"
#include "stk_rng.cuh"

#ifndef STK_GEMM_EPI_WARP0
#define STK_GEMM_EPI_WARP0 0
#endif

namespace stk {

constexpr int BM = 128, BN = 256, BK = 64;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_BUF_BYTES = 128 * 128;    // [128 rows][128 B] staging tile

// Kernel geometry per (epilogue, pairing).
// PAIR: two CTAs of a cluster (ranks 2p, 2p+1) run one 256 x 256 tile with cta_group::2 MMAs: each CTA
// stages its 128 rows of A and 128 of the 256 rows of B (32 KB per stage instead of 48 KB), the even CTA
// issues the MMAs for both, each CTA drains its own 128 x 256 accumulator.  Halving the B traffic per SM
// takes the operand reads + TMA fills under the shared-memory bandwidth that limits the single-CTA form.
// LN: the fused residual + LayerNorm epilogue (STK_EPI_BIAS_RESID_LN) spreads the 768-wide row over three
// column slabs (three CTAs, or three pairs) of one cluster, with two extra I/O warps that move the
// residual in and the results out through FOUR staging tiles by TMA.
template <int EPI, bool PAIR>
struct GemmCfg {
  static constexpr bool kLN = EPI == STK_EPI_BIAS_RESID_LN || EPI == STK_EPI_BIAS_DROP_RESID_LN;
  static constexpr int kBRows = PAIR ? 128 : 256;                   // rows of B this CTA stages
  static constexpr int kBStageBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = A_STAGE_BYTES + kBStageBytes;
  static constexpr int kStages = PAIR ? (kLN ? 4 : 6) : (kLN ? 3 : 4);
  static constexpr int kThreads = kLN ? 384 : 320;
  static constexpr int kEpiBufs = kLN ? 4 : 2;
  static constexpr int kStatsBytes = kLN ? 2 * 6 * 128 * 8 : 0;   // [parity][slab half][row] (mean, M2)
  static constexpr int kParamBytes = kLN ? 3 * 256 * 4 : 256 * 4;  // bias (+ gamma, beta) of this CTA's columns
  static constexpr int kCluster = (kLN ? 3 : 1) * (PAIR ? 2 : 1);
  static constexpr int kTileM = PAIR ? 256 : 128;
  static constexpr int kSmem = 1024 /*align slack*/ + kStages * kStageBytes + kEpiBufs * EPI_BUF_BYTES + kStatsBytes +
                               kParamBytes + 512 /*barriers*/;
};

// Warp roles.  The eight epilogue warps come FIRST (warp w drains TMEM lane quarter w % 4, column half w / 4),
// the TMA producer and the MMA issuer after them.
constexpr int kEpiWarp0 = STK_GEMM_EPI_WARP0;
constexpr int kProducerWarp = kEpiWarp0 == 0 ? 8 : 0;
constexpr int kMmaWarp = kEpiWarp0 == 0 ? 9 : 1;

// Bring-up instrumentation (clock64 timelines, "skip the loads / MMAs / epilogue math" switches driven by the env
// variable STK_GEMM_DEBUG) is compiled in only with -DSTK_GEMM_DEBUG_BUILD=1 (python -m stonkgs_b200.build --debug);
// the production instantiations carry none of it.
#ifndef STK_GEMM_DEBUG_BUILD
#define STK_GEMM_DEBUG_BUILD 0
#endif
#define STK_DBG(expr) (STK_GEMM_DEBUG_BUILD && (expr))

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  int dbg;   // bring-up only (env STK_GEMM_DEBUG): CTA 0 records a clock64 timeline of its first tiles
  StkGemmEpilogue epi;
};

__device__ long long g_gemm_timeline[4096];   // bring-up only: [tile][16] clock64 stamps of CTA 0
#define STK_GEMM_STAMP(cond, t, slot) \
  do { if (STK_DBG(p.dbg) && blockIdx.x < 3 && (cond) && (t) < 64) g_gemm_timeline[blockIdx.x * 1024 + (t) * 16 + (slot)] = clock64(); } while (0)

// ------------------------------------------------------------------------------------------------
// epilogue helpers
// ------------------------------------------------------------------------------------------------
// Write this thread's 128 B of row data into the swizzled staging tile, then have one thread of the
// 128-thread group issue the TMA store (or reduce-add) of the [128 x 128B] box at (c0, c1).
template <bool kReduceAdd, bool kAcquired = false>
__device__ __forceinline__ void stage_and_store(const CUtensorMap* map, uint8_t* buf, int row, const uint4 (&data)[8],
                                                int c0, int c1, bool store_thread, uint32_t bar_id) {
  if (!kAcquired) {
    if (store_thread) tma_wait_group_read<0>();  // previous store out of this buffer has been read
    named_bar_sync(bar_id, 128);
  }
  uint8_t* rowp = buf + row * 128;
#pragma unroll
  for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(rowp + ((c ^ (row & 7)) << 4)) = data[c];
  fence_proxy_async_smem();
  named_bar_sync(bar_id, 128);
  if (store_thread) {
    if (kReduceAdd)
      tma_reduce_add_2d(map, buf, c0, c1);
    else
      tma_store_2d(map, buf, c0, c1);
    tma_commit_group();
  }
}

// DYN: dynamic tile scheduling through cluster launch control.  The kernel is then launched with one CTA (pair) per work
// item; the CTAs that get an SM behave like persistent ones, but instead of walking a static stride they CANCEL a
// not-yet-launched CTA (pair) of the same grid and take over its item.  A CTA that cannot get an SM — because a
// collective's CTA sits there — therefore costs nothing: its item is simply stolen by the others (with the static
// schedule the kernel takes ~2 x as long).  Response ring of four 16-byte slots: the producer warp of the
// issuing CTA requests item k+1 when it starts item k; every role that walks the item sequence (producer, MMA issuer,
// eight epilogue warps) waits for the slot, decodes it and releases it.  Not used by the LayerNorm variants.
constexpr int kClcDepth = 4;

template <int A_MN, int B_MN, int EPI, bool PAIR, bool DYN>
__global__ void __launch_bounds__(GemmCfg<EPI, PAIR>::kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c2,
            const __grid_constant__ CUtensorMap map_r, const GemmParams p) {
  using Cfg = GemmCfg<EPI, PAIR>;
  constexpr bool kLN = Cfg::kLN;
  constexpr int STAGES = Cfg::kStages;
  constexpr int B_STAGE_BYTES = Cfg::kBStageBytes;
  constexpr int TM = Cfg::kTileM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_epi = smem_b + STAGES * B_STAGE_BYTES;
  float2* s_stats = reinterpret_cast<float2*>(smem_epi + Cfg::kEpiBufs * EPI_BUF_BYTES);   // LN only
  float* s_par = reinterpret_cast<float*>(smem_epi + Cfg::kEpiBufs * EPI_BUF_BYTES + Cfg::kStatsBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_par) + Cfg::kParamBytes);
  uint64_t* full_bar = bars;                 // [STAGES]  (PAIR: only the even CTA's are used)
  uint64_t* empty_bar = bars + 8;            // [STAGES]
  uint64_t* tfull_bar = bars + 16;           // [2]
  uint64_t* tempty_bar = bars + 18;          // [2]       (PAIR: only the even CTA's are used)
  // LN epilogue only:
  uint64_t* rfull_bar = bars + 20;           // [4] residual chunk landed in staging tile L
  uint64_t* staged_bar = bars + 24;          // [4] normalised output written to staging tile L
  uint64_t* zstaged_bar = bars + 28;         // [4] pre-LN sum written to staging tile L (training: saved for backward)
  uint64_t* zdone_bar = bars + 32;           // [4] ... and read out by its TMA store
  uint64_t* stats_bar = bars + 36;           // [2] row statistics of all three slabs have arrived (cluster scope)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 38);
  // DYN only:
  uint64_t* clc_full = bars + 40;            // [4] response k has landed in slot k % 4
  uint64_t* clc_empty = bars + 44;           // [4] every reader has released the slot (PAIR: the even CTA's are used)
  uint4* clc_resp = reinterpret_cast<uint4*>(bars + 48);   // [4] 16-byte try_cancel responses
  static_assert(STAGES <= 8, "barrier layout");
  static_assert(!(DYN && kLN), "the LayerNorm variants keep the static schedule");
  // readers of every response: producer warp + eight epilogue warps of each CTA, MMA warp of the issuing CTA
  constexpr int kClcReaders = PAIR ? 19 : 10;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = Cfg::kCluster > 1 ? cluster_ctarank() : 0u;   // rank in the cluster
  const uint32_t pr = PAIR ? (crank & 1u) : 0u;                        // rank in the CTA pair (0 issues the MMAs)
  const uint32_t pair_leader = crank & ~1u;
  // persistent scheduling unit = CTA or CTA pair
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int units = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == kProducerWarp && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (EPI != STK_EPI_CE_STATS) tma_prefetch_desc(&map_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, PAIR ? 16 : 8);  // one arrive per epilogue warp (of both CTAs of a pair)
    }
    if (DYN) {
      for (int i = 0; i < kClcDepth; ++i) {
        mbar_init(clc_full + i, 1);
        mbar_init(clc_empty + i, kClcReaders);
      }
    }
    if (kLN) {
      for (int i = 0; i < 4; ++i) {
        mbar_init(rfull_bar + i, 1);
        mbar_init(staged_bar + i, 4);    // one arrive per warp of the 128-thread group
        mbar_init(zstaged_bar + i, 4);
        mbar_init(zdone_bar + i, 1);
      }
      for (int i = 0; i < 2; ++i) mbar_init(stats_bar + i, 1);   // one local arrive.expect_tx + 6 x 128 x 8 B of st.async
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (PAIR) tmem_alloc_pair(tmem_slot, 512);
    else tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (Cfg::kCluster > 1) cluster_sync_all();   // peers' barriers are initialised before anyone arrives on them remotely
  const uint32_t tmem_base = *tmem_slot;

  const int num_items = p.m_tiles * p.n_tiles * p.splits;
  // item walk: static stride over the persistent grid, or (DYN) this CTA's own block index first and then whatever
  // try_cancel hands over.  clc_next(k): the item after the k-th one (-1 = nothing left); called by whole warps.
  const int first_item = DYN ? static_cast<int>(blockIdx.x) / Cfg::kCluster : unit;
  auto clc_next = [&](int k) -> int {
    const int sl = k % kClcDepth;
    mbar_wait(clc_full + sl, (k / kClcDepth) & 1);
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(smem_u32(clc_resp + sl)) : "memory");
    const int bx = clc_decode(r);
    __syncwarp();
    if (lane == 0) {
      if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(clc_empty + sl), pair_leader));
      else mbar_arrive(clc_empty + sl);
    }
    return bx < 0 ? -1 : bx / Cfg::kCluster;
  };
  auto next_item = [&](int item, int k) -> int {
    if (DYN) return clc_next(k);
    return item + units < num_items ? item + units : -1;
  };

  if (warp == kProducerWarp) {
    // ============================== TMA producer ==============================
    // Warp-uniform control flow (all lanes walk the loop and poll the barriers, so addresses and
    // coordinates stay on the uniform datapath); one elected lane issues the TMA instructions.
    const bool leader = elect_one();
    int stage = 0, n_loaded = 0;
    uint32_t phase = 0;
    int it_k = 0;
    for (int item = first_item; item >= 0 && item < num_items; item = next_item(item, it_k), ++it_k) {
      if (DYN) {
        // request the item that follows this one: the slot's previous response (k - 4) must have been released by all
        // of its readers; every CTA arms its own barrier, the even CTA of a pair issues for both (multicast)
        const int sl = it_k % kClcDepth;
        if (it_k >= kClcDepth && (!PAIR || pr == 0)) mbar_wait(clc_empty + sl, ((it_k / kClcDepth) - 1) & 1);
        if (leader) {
          mbar_arrive_expect_tx(clc_full + sl, 16);
          if (!PAIR) clc_try_cancel(smem_u32(clc_resp + sl), smem_u32(clc_full + sl));
          else if (pr == 0) clc_try_cancel_multicast(smem_u32(clc_resp + sl), smem_u32(clc_full + sl));
        }
        __syncwarp();
      }
      const int split = item % p.splits;
      const int tile = item / p.splits;
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (tile / p.n_tiles) * TM + static_cast<int>(pr) * 128;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        if (kb == kb0 && !kLN) STK_GEMM_STAMP(leader, it_k, 12);
        if (STK_DBG(p.dbg & 4) && n_loaded >= STAGES) {   // bring-up: no loads after the ring's first fill (pure MMA rate)
          if (leader && (!PAIR || pr == 0)) mbar_arrive(full_bar + stage);
        } else if (leader) {
          // PAIR: both CTAs' loads complete on the even CTA's barrier, which expects the bytes of both
          if (!PAIR || pr == 0) mbar_arrive_expect_tx(full_bar + stage, (PAIR ? 2 : 1) * Cfg::kStageBytes);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k0 = kb * BK;
          const int nb0 = n0 + static_cast<int>(pr) * 128;   // PAIR: this CTA stages rows [nb0, nb0 + 128) of B
          auto load = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
            if (PAIR) tma_load_2d_pair(m, full_bar + stage, dst, c0, c1);
            else tma_load_2d(m, full_bar + stage, dst, c0, c1);
          };
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) load(&map_a, sa + i * 8192, m0 + 64 * i, k0);
          } else {
            load(&map_a, sa, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < Cfg::kBRows / 64; ++i) load(&map_b, sb + i * 8192, nb0 + 64 * i, k0);
          } else {
            load(&map_b, sb, k0, nb0);
          }
        }
        __syncwarp();
        ++n_loaded;
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    if (PAIR && pr != 0) {
      // the odd CTA of a pair only lends its shared and tensor memory to the MMAs issued by the even one
    } else {
    // ============================== MMA issuer ==============================
    // Same structure: the whole warp tracks the pipeline state, one elected lane issues tcgen05.mma /
    // tcgen05.commit, so descriptor arithmetic is uniform and no per-instruction R2UR chain forms.
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(TM, BN, A_MN, B_MN);   // PAIR: M = 256 across the two CTAs
    const uint32_t pair_mask = 3u << pair_leader;
    // K-major: 8-row groups 1024 B apart, +32 B per 16-wide k step.
    // MN-major: 64-wide chunks 8192 B apart (LBO), 8 k-rows 1024 B apart (SBO), +2048 B per k step.
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), A_MN ? 8192 : 16, 1024);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), B_MN ? 8192 : 16, 1024);
    constexpr uint32_t a_kstep = (A_MN ? 2048 : 32) >> 4;
    constexpr uint32_t b_kstep = (B_MN ? 2048 : 32) >> 4;
    const uint32_t tmem_base_u = __reduce_max_sync(0xffffffffu, tmem_base);
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t as_phase = 0;
    bool ready = false;
    int it_k = 0;
    for (int item = first_item; item >= 0 && item < num_items; item = next_item(item, it_k), ++it_k) {
      const int split = item % p.splits;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
      mbar_wait(tempty_bar + as, as_phase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      // warp-uniform by construction (redux result): a per-lane register here makes the compiler wrap every
      // tcgen05.mma in an elect / R2UR.BROADCAST / branch "waterfall" (~180 cycles per MMA instead of 128)
      const uint32_t d_tmem = tmem_base_u + as * BN;
      STK_GEMM_STAMP(leader, it_k, 0);
      long long waited = 0;
      for (int kb = kb0; kb < kb1; ++kb) {
        // `ready` = the look-ahead probe issued between the previous k-block's MMAs already saw this slot full
        if (!ready) {
          if (STK_DBG(p.dbg)) {
            const long long t0 = clock64();
            mbar_wait(full_bar + stage, phase);
            waited += clock64() - t0;
          } else {
            mbar_wait(full_bar + stage, phase);
          }
        }
        tc_fence_after();
        const long long kb_t0 = STK_DBG(p.dbg) ? clock64() : 0;
        if (kb == kb0) STK_GEMM_STAMP(leader, it_k, 13);
        if (kb == kb1 - 1) STK_GEMM_STAMP(leader, it_k, 1);
        {
          // The whole (converged) warp executes every tcgen05 statement; elect.sync inside the statement picks
          // the issuing lane, all operands are warp-uniform: ptxas emits ELECT + @P UTCHMMA back to back and
          // the tensor pipe is fed at its 128-cycle cadence (a branch on a cached elect result costs ~180
          // cycles per MMA in elect / broadcast / loop overhead — measured with tools/mma_rate.py).
          const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
          const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
          auto commit = [&](uint64_t* bar) {
            if (PAIR) umma_commit_pair_warp(bar, pair_mask);
            else umma_commit_warp(bar);
          };
          // The tensor pipe buffers about one pending MMA: whatever the warp does between two issues is
          // hidden only while the previous MMA (128 cycles) executes.  So the NEXT slot's full barrier is
          // probed (non-blocking) in the middle of this k-block's issues, and the whole k-block — elect,
          // four MMAs, probe, slot-free commit — is one statement (umma_kblock_warp).
          const int nstage = stage + 1 == STAGES ? 0 : stage + 1;
          const uint32_t nphase = stage + 1 == STAGES ? phase ^ 1 : phase;
          static_assert(BK / 16 == 4, "k-steps per stage");
          if (!STK_DBG(p.dbg & 8)) {   // bring-up bit 8: no MMAs (pure TMA fill rate)
            ready = umma_kblock_warp<PAIR>(d_tmem, a_desc, b_desc, a_desc + a_kstep, b_desc + b_kstep,
                                           a_desc + 2 * a_kstep, b_desc + 2 * b_kstep, a_desc + 3 * a_kstep,
                                           b_desc + 3 * b_kstep, idesc, kb != kb0 ? 1u : 0u, empty_bar + stage, pair_mask,
                                           full_bar + nstage, nphase);
          } else {
            ready = false;
            commit(empty_bar + stage);
          }
          if (kb == kb1 - 1) commit(tfull_bar + as);
          if (STK_DBG(p.dbg) && blockIdx.x == 0 && leader && it_k == 2 && kb - kb0 < 64) {
            g_gemm_timeline[3072 + (kb - kb0) * 2] = kb_t0;          // k-block operands ready
            g_gemm_timeline[3072 + (kb - kb0) * 2 + 1] = clock64();  // its MMAs + commits issued
          }
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (STK_DBG(p.dbg) && blockIdx.x < 3 && leader && it_k < 64)
        g_gemm_timeline[blockIdx.x * 1024 + it_k * 16 + 14] = waited;   // cycles starved for operands
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
    }
  } else if (kLN && warp >= 10) {
    // ============================== epilogue I/O warps (LN variant) ==============================
    // Warp 10 + g serves column half g of this CTA's 128 x 256 slab through its two staging tiles
    // (64-column chunks): residual chunk in by TMA, (training: pre-LN sum out,) normalised rows out, next
    // tile's residual in.  The epilogue math warps never touch global memory for these tensors.
    if (lane == 0) {
      const int g = warp - 10;
      const int L0 = g * 2, L1 = g * 2 + 1;
      uint8_t* ebuf0 = smem_epi + L0 * EPI_BUF_BYTES;
      uint8_t* ebuf1 = smem_epi + L1 * EPI_BUF_BYTES;
      const bool save_z = p.epi.c2 != nullptr;
      uint32_t it = 0;
      for (int item = unit; item < num_items; item += units, ++it) {
        const int n_col = (item % p.n_tiles) * BN + g * 128;
        const int m0 = (item / p.n_tiles) * TM + static_cast<int>(pr) * 128;
        if (it == 0) {
          mbar_arrive_expect_tx(rfull_bar + L0, EPI_BUF_BYTES);
          tma_load_2d(&map_r, rfull_bar + L0, ebuf0, n_col, m0);
          mbar_arrive_expect_tx(rfull_bar + L1, EPI_BUF_BYTES);
          tma_load_2d(&map_r, rfull_bar + L1, ebuf1, n_col + 64, m0);
        }
        if (save_z) {
          mbar_wait(zstaged_bar + L0, it & 1);
          tma_store_2d(&map_c2, ebuf0, n_col, m0);
          tma_commit_group();
          mbar_wait(zstaged_bar + L1, it & 1);
          tma_store_2d(&map_c2, ebuf1, n_col + 64, m0);
          tma_commit_group();
          tma_wait_group_read<1>();
          mbar_arrive(zdone_bar + L0);
          tma_wait_group_read<0>();
          mbar_arrive(zdone_bar + L1);
        }
        mbar_wait(staged_bar + L0, it & 1);
        tma_store_2d(&map_c, ebuf0, n_col, m0);
        tma_commit_group();
        mbar_wait(staged_bar + L1, it & 1);
        tma_store_2d(&map_c, ebuf1, n_col + 64, m0);
        tma_commit_group();
        STK_GEMM_STAMP(g == 0, static_cast<int>(it), 10);
        const int next = item + units;
        const int m0n = (next / p.n_tiles) * TM + static_cast<int>(pr) * 128;
        tma_wait_group_read<1>();
        if (next < num_items) {
          mbar_arrive_expect_tx(rfull_bar + L0, EPI_BUF_BYTES);
          tma_load_2d(&map_r, rfull_bar + L0, ebuf0, n_col, m0n);
        }
        tma_wait_group_read<0>();
        if (next < num_items) {
          mbar_arrive_expect_tx(rfull_bar + L1, EPI_BUF_BYTES);
          tma_load_2d(&map_r, rfull_bar + L1, ebuf1, n_col + 64, m0n);
        }
        STK_GEMM_STAMP(g == 0, static_cast<int>(it), 15);
      }
      tma_wait_group<0>();
    }
    __syncwarp();
  } else if (kLN) {
    // ============================== epilogue warps, fused bias + residual + LayerNorm ==============================
    // HF BertSelfOutput / BertOutput (modeling_bert.py:294-298, 352-356): y = LN(dense(x) + residual), eps 1e-12.
    // The 768-wide row is spread over the three CTAs of the cluster.  Every thread owns one row x 128
    // columns: pass 1 forms z = acc + bias + residual and its (mean, M2) over those columns; the six
    // partials of a row are exchanged through distributed shared memory (every thread pushes its partial
    // into all three CTAs with st.async, which completes transaction bytes on the receiver's mbarrier:
    // no fences), merged with the parallel-variance formula, and pass 2 normalises the bf16 z kept in
    // registers.
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int g = ew >> 2;    // column half of the 256-wide accumulator
    const int row = q * 32 + lane;
    const StkGemmEpilogue& e = p.epi;
    const uint32_t rank = PAIR ? crank >> 1 : crank;   // column slab == n-tile index of this CTA (grid is a multiple of the cluster size)
    {
      const int t = ew * 32 + lane;            // 0..255: column of this CTA's slab
      const int n = static_cast<int>(rank) * BN + t;
      s_par[t] = e.bias ? __ldg(e.bias + n) : 0.f;
      s_par[256 + t] = __ldg(e.ln_gamma + n);
      s_par[512 + t] = __ldg(e.ln_beta + n);
    }
    named_bar_sync(3, 256);
    const bool save_z = e.c2 != nullptr;
    const uint32_t stats_bar_addr[2] = {smem_u32(stats_bar), smem_u32(stats_bar + 1)};
    // train(): z = drop(acc + bias) + residual (HF:296-298, 354-356).  The keep decision is the pure function of
    // (seed, site, token row, hidden column) of stk_rng.cuh — the same one the LayerNorm backward regenerates.
    constexpr bool kDrop = EPI == STK_EPI_BIAS_DROP_RESID_LN;
    const uint32_t drop_thr4 = kDrop ? drop_thr4_of(e.drop_thr) : 0u;
    const f32x2_t drop_scale2 = kDrop ? pack_f32x2(drop_scale(e.drop_thr), drop_scale(e.drop_thr)) : 0ull;

    int as = 0;
    uint32_t as_phase = 0, it = 0;
    for (int item = unit; item < num_items; item += units, ++it) {
      const int m0 = (item / p.n_tiles) * TM + static_cast<int>(pr) * 128;
      const int m = m0 + row;
      const int dbg_t = static_cast<int>(it);
      const bool dbg_thr = threadIdx.x == kEpiWarp0 * 32;
      STK_GEMM_STAMP(dbg_thr, dbg_t, 2);
      mbar_wait(tfull_bar + as, as_phase);
      tc_fence_after();
      STK_GEMM_STAMP(dbg_thr, dbg_t, 3);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + g * 128;

      uint4 zpk[2][8];
      float cmean[2], cm2[2];
      const uint32_t drop_key = kDrop ? drop_row_key(e.drop_seed, e.drop_site, static_cast<uint32_t>(m)) : 0u;
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(t_row + chunk * 64, r[0]);
        tmem_ld_32x32b_x32(t_row + chunk * 64 + 32, r[1]);
        tmem_ld_wait();
        if (chunk == 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(tempty_bar + as), pair_leader));
            else mbar_arrive(tempty_bar + as);
          }
        }
        const int L = g * 2 + chunk;
        uint8_t* ebuf = smem_epi + L * EPI_BUF_BYTES;
        mbar_wait(rfull_bar + L, it & 1);
        STK_GEMM_STAMP(dbg_thr, dbg_t, 4 + chunk * 4);
        const float4* bias4 = reinterpret_cast<const float4*>(s_par + g * 128 + chunk * 64);
        // packed fp32x2 math: v[k] = (z[2k], z[2k+1])
        f32x2_t v[32];
        f32x2_t s2 = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 ex4 = *reinterpret_cast<const uint4*>(ebuf + row * 128 + ((c ^ (row & 7)) << 4));
          const uint32_t ex[4] = {ex4.x, ex4.y, ex4.z, ex4.w};
          const float4 b0 = bias4[2 * c], b1 = bias4[2 * c + 1];
          const f32x2_t bv[4] = {pack_f32x2(b0.x, b0.y), pack_f32x2(b0.z, b0.w), pack_f32x2(b1.x, b1.y),
                                 pack_f32x2(b1.z, b1.w)};
          f32x2_t keep[4];   // kDrop: all-ones / all-zeros per element of the four column pairs of this 8-column group
          if (kDrop) {
            uint32_t w0, w1;
            drop_words(drop_key, static_cast<uint32_t>((static_cast<int>(rank) * BN + g * 128 + chunk * 64 + c * 8) >> 3), w0, w1);
            const uint32_t sg0 = drop_signs(w0, drop_thr4), sg1 = drop_signs(w1, drop_thr4);
            keep[0] = pack_u32x2(drop_mask32<0>(sg0), drop_mask32<1>(sg0));
            keep[1] = pack_u32x2(drop_mask32<2>(sg0), drop_mask32<3>(sg0));
            keep[2] = pack_u32x2(drop_mask32<0>(sg1), drop_mask32<1>(sg1));
            keep[3] = pack_u32x2(drop_mask32<2>(sg1), drop_mask32<3>(sg1));
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = c * 8 + i * 2;
            const f32x2_t acc = pack_f32x2(__uint_as_float(r[j >> 5][j & 31]), __uint_as_float(r[(j + 1) >> 5][(j + 1) & 31]));
            f32x2_t z2;
            if (kDrop) z2 = fma_f32x2(add_f32x2(acc, bv[i]) & keep[i], drop_scale2, bf16x2_to_f32x2(ex[i]));
            else z2 = add_f32x2(add_f32x2(acc, bv[i]), bf16x2_to_f32x2(ex[i]));
            v[c * 4 + i] = z2;
            s2 = add_f32x2(s2, z2);
          }
        }
        float sa, sb;
        unpack_f32x2(s2, sa, sb);
        const float mu = (sa + sb) * (1.0f / 64.0f);
        const f32x2_t nmu2 = pack_f32x2(-mu, -mu);
        f32x2_t q2 = pack_f32x2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const f32x2_t d2 = add_f32x2(v[c * 4 + i], nmu2);
            q2 = fma_f32x2(d2, d2, q2);
            w[i] = f32x2_to_bf16x2(v[c * 4 + i]);
          }
          zpk[chunk][c] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        float q0, q1;
        unpack_f32x2(q2, q0, q1);
        cmean[chunk] = mu;
        cm2[chunk] = q0 + q1;
        if (save_z) {   // training: z is kept for the LayerNorm backward (second output, row pitch ldc2)
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(ebuf + row * 128 + ((c ^ (row & 7)) << 4)) = zpk[chunk][c];
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(zstaged_bar + L);
        }
        STK_GEMM_STAMP(dbg_thr, dbg_t, 5 + chunk * 4);
      }

      // ---- row statistics: merge the two 64-column chunks, publish to the cluster, merge six partials ----
      const uint32_t par = it & 1;
      {
        const float d = cmean[1] - cmean[0];
        const float mean_t = 0.5f * (cmean[0] + cmean[1]);
        const float m2_t = cm2[0] + cm2[1] + d * d * 32.0f;   // n_a n_b / (n_a + n_b) = 32
        const uint32_t slot = smem_u32(s_stats + (par * 6 + rank * 2 + g) * 128 + row);
        if (ew == 0 && lane == 0) mbar_arrive_expect_tx(stats_bar + par, 6 * 128 * 8);
#pragma unroll
        for (uint32_t k = 0; k < 3; ++k) {   // the CTAs holding the other slabs of the same rows (and this one)
          const uint32_t peer = PAIR ? 2 * k + pr : k;
          st_async_f32x2(map_to_cta(slot, peer), mean_t, m2_t, map_to_cta(stats_bar_addr[par], peer));
        }
      }
      STK_GEMM_STAMP(dbg_thr, dbg_t, 12);
      mbar_wait(stats_bar + par, (it >> 1) & 1);
      STK_GEMM_STAMP(dbg_thr, dbg_t, 6);
      float mean, rstd;
      {
        float2 pt[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) pt[k] = s_stats[(par * 6 + k) * 128 + row];
        float sm = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) sm += pt[k].x;
        mean = sm * (1.0f / 6.0f);
        float m2 = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          const float d = pt[k].x - mean;
          m2 += pt[k].y + 128.0f * d * d;
        }
        rstd = rsqrtf(m2 * (1.0f / 768.0f) + kLnEps);
      }
      if (e.ln_mean != nullptr && rank == 0 && g == 0 && m < p.M) {
        e.ln_mean[m] = mean;
        e.ln_rstd[m] = rstd;
      }

      // ---- pass 2: normalise, scale, shift; hand the staging tile to the I/O warp ----
#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {
        const int L = g * 2 + chunk;
        uint8_t* ebuf = smem_epi + L * EPI_BUF_BYTES;
        const float4* gam4 = reinterpret_cast<const float4*>(s_par + 256 + g * 128 + chunk * 64);
        const float4* bet4 = reinterpret_cast<const float4*>(s_par + 512 + g * 128 + chunk * 64);
        if (save_z) mbar_wait(zdone_bar + L, it & 1);   // the z store has read the tile: it may be overwritten
        const f32x2_t rstd2 = pack_f32x2(rstd, rstd);
        const f32x2_t nmr2 = pack_f32x2(-mean * rstd, -mean * rstd);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t zz[4] = {zpk[chunk][c].x, zpk[chunk][c].y, zpk[chunk][c].z, zpk[chunk][c].w};
          const float4 g0 = gam4[2 * c], g1 = gam4[2 * c + 1];
          const float4 h0 = bet4[2 * c], h1 = bet4[2 * c + 1];
          const f32x2_t gv[4] = {pack_f32x2(g0.x, g0.y), pack_f32x2(g0.z, g0.w), pack_f32x2(g1.x, g1.y),
                                 pack_f32x2(g1.z, g1.w)};
          const f32x2_t hv[4] = {pack_f32x2(h0.x, h0.y), pack_f32x2(h0.z, h0.w), pack_f32x2(h1.x, h1.y),
                                 pack_f32x2(h1.z, h1.w)};
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const f32x2_t t2 = fma_f32x2(bf16x2_to_f32x2(zz[i]), rstd2, nmr2);   // (z - mean) * rstd
            w[i] = f32x2_to_bf16x2(fma_f32x2(t2, gv[i], hv[i]));
          }
          *reinterpret_cast<uint4*>(ebuf + row * 128 + ((c ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(staged_bar + L);
        STK_GEMM_STAMP(dbg_thr, dbg_t, 7 + chunk * 4);
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else {
    // ============================== epilogue warps ==============================
    const int ew = warp - kEpiWarp0;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int g = ew >> 2;    // column half of the 256-wide accumulator
    const int row = q * 32 + lane;
    uint8_t* buf = smem_epi + g * EPI_BUF_BYTES;
    const bool store_thread = ((ew & 3) == 0) && lane == 0;
    const uint32_t bar_id = 1 + g;
    const StkGemmEpilogue& e = p.epi;
    float ce_scale = 0.f;
    if (EPI == STK_EPI_CE_DLOGIT) ce_scale = __ldg(e.scale_dev);
    constexpr bool kHasBias = EPI == STK_EPI_BIAS || EPI == STK_EPI_BIAS_GELU || EPI == STK_EPI_BIAS_GELU_SAVE ||
                              EPI == STK_EPI_BIAS_GELU_SAVE_GRAD || EPI == STK_EPI_BIAS_RESID || EPI == STK_EPI_BIAS_TANH_F32;
    constexpr bool kSaves2 = EPI == STK_EPI_BIAS_GELU_SAVE || EPI == STK_EPI_BIAS_GELU_SAVE_GRAD;   // second bf16 output
    const int gt = (ew & 3) * 32 + lane;   // thread index within the epilogue group
    float* s_bias = s_par + g * 128;       // bias of this group's 128 columns of the current tile (0 beyond N)

    int as = 0;
    uint32_t as_phase = 0;
    int it_k = 0;
    for (int item = first_item; item >= 0 && item < num_items; item = next_item(item, it_k), ++it_k) {
      const int tile = item / p.splits;
      const int n_blk = tile % p.n_tiles;
      const int n0 = n_blk * BN;
      const int m0 = (tile / p.n_tiles) * TM + static_cast<int>(pr) * 128;
      const int m = m0 + row;
      const bool m_ok = m < p.M;
      const int dbg_t = it_k;
      const bool dbg_thr = threadIdx.x == kEpiWarp0 * 32;
      STK_GEMM_STAMP(dbg_thr, dbg_t, 2);
      if (kHasBias) {
        // every thread of the previous tile is past its last read of s_bias (the staging barriers of
        // its final chunk come after the math), so the vector can be replaced now
        const int n = n0 + g * 128 + gt;
        s_bias[gt] = (e.bias != nullptr && n < p.N) ? __ldg(e.bias + n) : 0.f;
        named_bar_sync(bar_id, 128);
      }
      // Residual / saved pre-activation rows do not depend on the accumulator: fetch both 64-column
      // chunks of this thread's row now so the global-load latency hides behind the MMA of this tile.
      constexpr bool kPrefetchExtra = EPI == STK_EPI_BIAS_RESID || EPI == STK_EPI_DGELU || EPI == STK_EPI_MUL;
      // COALESCED: thread t of the 128-thread group fetches 16-byte piece (t + 128 i) of the
      // [128 rows x 128 B] residual tile (8 consecutive threads = one full 128-byte row segment); the
      // pieces are transposed to the thread-per-row accumulator layout through the staging tile later.
      uint4 ex_pre[2][8];
      if (kPrefetchExtra) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const int ncp = n0 + g * 128 + ch * 64;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = gt + 128 * i;
            const int rr = idx >> 3, cc = idx & 7;
            const bool okp = (m0 + rr) < p.M && ncp + 64 <= p.N;
            const uint4* pp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(e.resid) +
                                                             static_cast<int64_t>(m0 + rr) * e.ldr + ncp) + cc;
            ex_pre[ch][i] = okp ? __ldg(pp) : make_uint4(0, 0, 0, 0);
          }
        }
      }
      mbar_wait(tfull_bar + as, as_phase);
      tc_fence_after();
      STK_GEMM_STAMP(dbg_thr, dbg_t, 3);
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + g * 128;

      // per-row CE state
      float ce_max = -INFINITY, ce_sum = 0.f;
      int label = -1;
      bool row_on = m_ok;   // CE: rows with a negative label are padding of a fixed-capacity row list (stk_compact_labels)
      float row_lse = 0.f;
      if (EPI == STK_EPI_CE_STATS || EPI == STK_EPI_CE_DLOGIT) {
        if (m_ok) {
          const int raw = __ldg(e.labels + m);
          row_on = raw >= 0;
          label = row_on ? raw - e.n_offset : -1;  // column within this call's B block
        }
        if (EPI == STK_EPI_CE_DLOGIT && m_ok) row_lse = __ldg(e.lse + m);
      }

#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {  // 64 accumulator columns per step
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(t_row + chunk * 64, r[0]);
        tmem_ld_32x32b_x32(t_row + chunk * 64 + 32, r[1]);
        tmem_ld_wait();
        STK_GEMM_STAMP(dbg_thr, dbg_t, 4 + chunk * 4);
        if (chunk == 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (PAIR) mbar_arrive_remote(map_to_cta(smem_u32(tempty_bar + as), pair_leader));
            else mbar_arrive(tempty_bar + as);
          }
        }
        const int nc = n0 + g * 128 + chunk * 64;  // first global column of this chunk
        if (STK_DBG(p.dbg & 16)) continue;   // bring-up: accumulator read only (isolates the epilogue's effect on the MMA rate)
        if (STK_DBG(p.dbg & 32) && q == 1 && EPI == STK_EPI_BIAS_GELU) {   // bring-up: no epilogue math on the MMA warp's scheduler
          uint4 data[8] = {};
          stage_and_store<false, false>(&map_c, buf, row, data, nc, m0, store_thread, bar_id);
          continue;
        }

        if (EPI == STK_EPI_CE_STATS) {
          float cmax = -INFINITY;
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int n = nc + h * 32 + j;
              float v = __uint_as_float(r[h][j]);
              v = (n < p.N) ? v : -INFINITY;
              r[h][j] = __float_as_uint(v);
              cmax = fmaxf(cmax, v);
              if (n == label && row_on) e.tgt_logit[m] = v;
            }
          const float new_max = fmaxf(ce_max, cmax);
          if (new_max > -INFINITY) {
            float s = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int j = 0; j < 32; ++j)
                s += fast_exp2((__uint_as_float(r[h][j]) - new_max) * 1.4426950408889634f);
            ce_sum = ce_sum * fast_exp2((ce_max - new_max) * 1.4426950408889634f) + s;
            ce_max = new_max;
          }
          continue;
        }

        // ---- elementwise epilogues: produce 128 B of output per thread and 32- or 64-column store ----
        if (EPI == STK_EPI_F32 || EPI == STK_EPI_F32_ADD || EPI == STK_EPI_BIAS_TANH_F32) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint4 data[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              float v[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[i] = __uint_as_float(r[h][c * 4 + i]);
                if (EPI == STK_EPI_BIAS_TANH_F32) v[i] = tanhf(v[i] + s_bias[chunk * 64 + h * 32 + c * 4 + i]);
              }
              data[c] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                                   __float_as_uint(v[3]));
            }
            stage_and_store<EPI == STK_EPI_F32_ADD>(&map_c, buf, row, data, nc + h * 32, m0, store_thread, bar_id);
          }
          continue;
        }

        // bf16 outputs
        constexpr bool kHasExtra = EPI == STK_EPI_BIAS_RESID || EPI == STK_EPI_DGELU || EPI == STK_EPI_MUL;
        if (kHasExtra) {
          // acquire the staging tile, drop the coalesced residual pieces into it (swizzled), then every
          // thread picks up its own row below; the result overwrites the same 16-byte slots
          if (store_thread) tma_wait_group_read<0>();
          named_bar_sync(bar_id, 128);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = gt + 128 * i;
            const int rr = idx >> 3, cc = idx & 7;
            *reinterpret_cast<uint4*>(buf + rr * 128 + ((cc ^ (rr & 7)) << 4)) = ex_pre[chunk][i];
          }
          named_bar_sync(bar_id, 128);
        }
        STK_GEMM_STAMP(dbg_thr, dbg_t, 5 + chunk * 4);
        uint4 data[8];
        uint4 data2[8];
        const float4* bias4 = reinterpret_cast<const float4*>(s_bias + chunk * 64);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4], w2[4];
          float bias_v[8];
          if (kHasBias) {   // broadcast shared-memory reads: no global loads, no tail predicates on the hot path
            const float4 b0 = bias4[2 * c], b1 = bias4[2 * c + 1];
            bias_v[0] = b0.x; bias_v[1] = b0.y; bias_v[2] = b0.z; bias_v[3] = b0.w;
            bias_v[4] = b1.x; bias_v[5] = b1.y; bias_v[6] = b1.z; bias_v[7] = b1.w;
          }
          uint4 ex4 = make_uint4(0, 0, 0, 0);
          if (kHasExtra) ex4 = *reinterpret_cast<const uint4*>(buf + row * 128 + ((c ^ (row & 7)) << 4));
          const uint32_t ex[4] = {ex4.x, ex4.y, ex4.z, ex4.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int j = c * 8 + i * 2;  // column within the 64-wide chunk
            float v0 = __uint_as_float(r[j >> 5][j & 31]);
            float v1 = __uint_as_float(r[(j + 1) >> 5][(j + 1) & 31]);
            if (kHasBias) {
              v0 += bias_v[i * 2];
              v1 += bias_v[i * 2 + 1];
            }
            if (EPI == STK_EPI_BIAS_GELU_SAVE) w2[i] = pack_bf16x2(v0, v1);
            if (EPI == STK_EPI_BIAS_GELU || EPI == STK_EPI_BIAS_GELU_SAVE) {
              v0 = gelu_erf(v0);
              v1 = gelu_erf(v1);
            }
            if (EPI == STK_EPI_BIAS_GELU_SAVE_GRAD) {
              float d0, d1;
              gelu_erf_with_grad(v0, v0, d0);
              gelu_erf_with_grad(v1, v1, d1);
              w2[i] = pack_bf16x2(d0, d1);
            }
            if (EPI == STK_EPI_MUL) {
              v0 *= bf16_lo(ex[i]);
              v1 *= bf16_hi(ex[i]);
            }
            if (EPI == STK_EPI_BIAS_RESID) {
              v0 += bf16_lo(ex[i]);
              v1 += bf16_hi(ex[i]);
            }
            if (EPI == STK_EPI_DGELU) {
              v0 *= gelu_erf_grad(bf16_lo(ex[i]));
              v1 *= gelu_erf_grad(bf16_hi(ex[i]));
            }
            if (EPI == STK_EPI_CE_DLOGIT) {
              const float l2e = 1.4426950408889634f;
              float p0 = fast_exp2((v0 - row_lse) * l2e);
              float p1 = fast_exp2((v1 - row_lse) * l2e);
              const int col = nc + j;  // column within this call's B block
              if (col == label) p0 -= 1.f;
              if (col + 1 == label) p1 -= 1.f;
              v0 = row_on ? p0 * ce_scale : 0.f;
              v1 = row_on ? p1 * ce_scale : 0.f;
            }
            w[i] = pack_bf16x2(v0, v1);
          }
          data[c] = make_uint4(w[0], w[1], w[2], w[3]);
          if (kSaves2) data2[c] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        }
        STK_GEMM_STAMP(dbg_thr, dbg_t, 6 + chunk * 4);
        stage_and_store<false, kHasExtra>(&map_c, buf, row, data, nc, m0, store_thread, bar_id);
        if (kSaves2)
          stage_and_store<false>(&map_c2, buf, row, data2, nc, m0, store_thread, bar_id);
        STK_GEMM_STAMP(dbg_thr, dbg_t, 7 + chunk * 4);
      }

      if (EPI == STK_EPI_CE_STATS && m_ok) {
        const int64_t slab = (e.n_offset >> 7) + n_blk * 2 + g;
        float2* out = reinterpret_cast<float2*>(e.ce_partial) + static_cast<int64_t>(m) * e.ce_pitch + slab;
        *out = make_float2(ce_max, ce_sum);
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
    if (store_thread) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (Cfg::kCluster > 1) cluster_sync_all();   // no CTA leaves while a peer may still touch its shared / tensor memory
  if (warp == kMmaWarp) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
extern std::atomic<long long> g_launches;

// 0 = static persistent schedule, 1 = dynamic (cluster launch control); stk_set_gemm_dynamic / env STK_GEMM_DYNAMIC
static std::atomic<int> g_gemm_dynamic{-1};

static bool gemm_dynamic() {
  int v = g_gemm_dynamic.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("STK_GEMM_DYNAMIC");
    v = e ? (atoi(e) != 0) : 0;
    g_gemm_dynamic.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

template <int A_MN, int B_MN, int EPI, bool PAIR, bool DYN>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mc2,
                  const CUtensorMap& mr, const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<EPI, PAIR>;
  auto kern = gemm_kernel<A_MN, B_MN, EPI, PAIR, DYN>;
  static bool configured[64] = {};
  static int max_clusters[64] = {};
  int dev = 0;
  STK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    STK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured[dev & 63] = true;
  }
  const int items = p.m_tiles * p.n_tiles * p.splits;
  if (Cfg::kCluster == 1) {
    const int sms = persistent_sms(dev);
    // DYN: one CTA per item; the ones that get an SM steal the rest
    kern<<<DYN ? items : (items < sms ? items : sms), Cfg::kThreads, Cfg::kSmem, stream>>>(ma, mb, mc, mc2, mr, p);
  } else {
    // clusters: CTA pairs (cta_group::2 MMAs), three column slabs of a LayerNorm row, or three pairs;
    // persistent over the tiles, as many clusters as the device can keep resident
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = Cfg::kCluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(Cfg::kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmem;
    cfg.stream = stream;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters[dev & 63] == 0) {
      cfg.gridDim = dim3(Cfg::kCluster * 32);
      int n = 0;
      STK_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      if (n <= 0) {
        set_error("stk_gemm: no %d-CTA cluster of this kernel fits on the device", Cfg::kCluster);
        return STK_ERR_UNSUPPORTED;
      }
      max_clusters[dev & 63] = n;
    }
    // work per cluster: a LayerNorm cluster takes whole row tiles (its slabs are the n-tiles), a pair one item
    const int work = Cfg::kLN ? p.m_tiles : items;
    const int reserved = num_sms(dev) - persistent_sms(dev);   // SMs left free for a concurrent collective
    int fit = max_clusters[dev & 63] - (reserved + Cfg::kCluster - 1) / Cfg::kCluster;
    if (fit < 1) fit = 1;
    const int clusters = DYN ? work : (work < fit ? work : fit);
    cfg.gridDim = dim3(Cfg::kCluster * clusters);
    STK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ma, mb, mc, mc2, mr, p));
  }
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

// Split-K factor of a reduce-add GEMM (split_k = 0).  Cost model in tensor-pipe cycles: the persistent grid runs
// ceil(tiles * s / units) waves (units = CTA pairs on 256 x 256 tiles when there is more than one 128-row tile, else
// single CTAs on 128 x 256), a wave costs the split's k-blocks (512 cycles each: four 128-cycle MMAs) plus a fixed
// per-item overhead (ring fill + the fp32 TMA reduce-add of the tile, ~8 k cycles).  The first version maximised grid
// occupancy alone and cut the cross-entropy backward's GEMMs into 23-k-block splits whose reduce-adds cost as much as
// their MMAs (dT at 525 TFLOP/s).
static int auto_splits(int M, int N, int K, int sms) {
  static int overhead = -1;
  if (overhead < 0) {
    const char* e = getenv("STK_SPLITK_OVERHEAD");
    overhead = e ? atoi(e) : 8000;
  }
  const bool pair = M > BM;
  const int tiles = ((M + (pair ? 2 * BM : BM) - 1) / (pair ? 2 * BM : BM)) * ((N + BN - 1) / BN);
  const int units = pair ? sms / 2 : sms;
  const int kb = (K + BK - 1) / BK;
  int best = 1;
  long long best_cost = -1;
  for (int s = 1; s <= (kb < 32 ? kb : 32); ++s) {
    const int per = (kb + s - 1) / s;
    const int real = (kb + per - 1) / per;   // splits that actually get k-blocks
    const long long waves = (static_cast<long long>(tiles) * real + units - 1) / units;
    const long long cost = waves * (static_cast<long long>(per) * 512 + overhead);
    if (best_cost < 0 || cost < best_cost) { best = s; best_cost = cost; }
  }
  return best;
}

}  // namespace stk

using namespace stk;

extern "C" int stk_gemm(int device, void* stream_, int a_major, int b_major, const void* A, int64_t lda,
                        const void* B, int64_t ldb, int M, int N, int K, int epilogue, void* C, int64_t ldc,
                        const StkGemmEpilogue* epi, int split_k) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  STK_REQUIRE(M > 0 && N > 0 && K > 0, "stk_gemm: empty problem M=%d N=%d K=%d", M, N, K);
  STK_REQUIRE(A && B, "stk_gemm: null operand");
  STK_REQUIRE((a_major | b_major) >> 1 == 0, "stk_gemm: major must be 0 or 1");
  STK_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "stk_gemm: lda/ldb must be multiples of 8 elements (16 B)");
  STK_REQUIRE((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(B) & 15) == 0,
              "stk_gemm: operands must be 16-byte aligned");
  // CTA pairs (256-row tiles, cta_group::2) whenever there is more than one 128-row tile of work
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("STK_GEMM_PAIR");
    pair_env = e ? atoi(e) : 1;
  }
  const bool pair = pair_env != 0 && M > BM;
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = pair ? (M + 2 * BM - 1) / (2 * BM) : (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.kb_total = (K + BK - 1) / BK;
  int splits = split_k < 0 ? 1 : split_k;
  if (epilogue != STK_EPI_F32_ADD) splits = 1;
  else if (splits == 0) splits = auto_splits(M, N, K, persistent_sms(device));
  if (splits > p.kb_total) splits = p.kb_total;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (epi) p.epi = *epi;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("STK_GEMM_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  p.dbg = dbg;

  const bool f32_out = epilogue == STK_EPI_F32 || epilogue == STK_EPI_F32_ADD || epilogue == STK_EPI_BIAS_TANH_F32;
  const bool has_c = epilogue != STK_EPI_CE_STATS;
  if (epilogue == STK_EPI_BIAS_RESID || epilogue == STK_EPI_DGELU || epilogue == STK_EPI_MUL) {
    STK_REQUIRE(epi && epi->resid && epi->ldr % 8 == 0 && N % 64 == 0, "stk_gemm: residual epilogue needs resid, ldr%%8==0, N%%64==0");
  }
  if (epilogue == STK_EPI_BIAS_GELU_SAVE || epilogue == STK_EPI_BIAS_GELU_SAVE_GRAD)
    STK_REQUIRE(epi && epi->c2 && epi->ldc2 % 8 == 0, "stk_gemm: GELU_SAVE needs c2");
  const bool ln_epi = epilogue == STK_EPI_BIAS_RESID_LN || epilogue == STK_EPI_BIAS_DROP_RESID_LN;
  if (epilogue == STK_EPI_BIAS_DROP_RESID_LN)
    STK_REQUIRE(epi && epi->drop_thr > 0 && epi->drop_thr < 128, "stk_gemm: the dropout LayerNorm epilogue needs 0 < drop_thr < 128");
  if (ln_epi) {
    STK_REQUIRE(N == kHidden && a_major == 0 && b_major == 0, "stk_gemm: the LayerNorm epilogue needs N == 768 and K-major operands");
    STK_REQUIRE(epi && epi->resid && epi->ldr % 8 == 0 && epi->ln_gamma && epi->ln_beta,
                "stk_gemm: the LayerNorm epilogue needs resid (ldr%%8==0), ln_gamma and ln_beta");
    STK_REQUIRE((reinterpret_cast<uintptr_t>(epi->resid) & 15) == 0, "stk_gemm: resid must be 16-byte aligned");
    STK_REQUIRE(epi->c2 == nullptr || (epi->ldc2 % 8 == 0 && (reinterpret_cast<uintptr_t>(epi->c2) & 15) == 0),
                "stk_gemm: c2 (pre-LayerNorm output) must be 16-byte aligned with ldc2%%8==0");
    STK_REQUIRE((epi->ln_mean == nullptr) == (epi->ln_rstd == nullptr), "stk_gemm: ln_mean and ln_rstd come together");
  }
  if (epilogue == STK_EPI_CE_STATS)
    STK_REQUIRE(epi && epi->labels && epi->ce_partial && epi->tgt_logit && epi->n_offset % 256 == 0, "stk_gemm: CE_STATS args");
  if (epilogue == STK_EPI_CE_DLOGIT)
    STK_REQUIRE(epi && epi->labels && epi->lse && epi->scale_dev && epi->n_offset % 256 == 0, "stk_gemm: CE_DLOGIT args");
  if (has_c) {
    STK_REQUIRE(C != nullptr && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "stk_gemm: C null or misaligned");
    STK_REQUIRE(ldc % (f32_out ? 4 : 8) == 0, "stk_gemm: ldc must be a multiple of 16 bytes");
  }

  STK_CHECK_CUDA(cudaSetDevice(device));   // after argument validation: bad arguments are reported without a device
  CUtensorMap ma, mb, mc, mc2, mr;
  int rc;
  // A: K-major -> stored [M][K], box {64 k, 128 m};  MN-major -> stored [K][M], box {64 m, 64 k}
  if (a_major == 0) rc = make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda * 2, 64, BM);
  else rc = make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda * 2, 64, 64);
  if (rc) return rc;
  if (b_major == 0) rc = make_tmap_2d(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, ldb * 2, 64, pair ? BN / 2 : BN);
  else rc = make_tmap_2d(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, N, K, ldb * 2, 64, 64);
  if (rc) return rc;
  if (has_c) {
    if (f32_out) rc = make_tmap_2d(&mc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, C, N, M, ldc * 4, 32, 128);
    else rc = make_tmap_2d(&mc, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, C, N, M, ldc * 2, 64, 128);
    if (rc) return rc;
  } else {
    mc = ma;
  }
  mc2 = mc;
  mr = mc;
  if (epilogue == STK_EPI_BIAS_GELU_SAVE || epilogue == STK_EPI_BIAS_GELU_SAVE_GRAD ||
      (ln_epi && epi->c2)) {
    rc = make_tmap_2d(&mc2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, epi->c2, N, M, epi->ldc2 * 2, 64, 128);
    if (rc) return rc;
  }
  if (ln_epi) {
    rc = make_tmap_2d(&mr, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, epi->resid, N, M, epi->ldr * 2, 64, 128);
    if (rc) return rc;
  }
  const bool dyn = gemm_dynamic() && !ln_epi;
#define STK_GEMM_CASE(AM, BMJ, E)                                                                                    \
  if (a_major == AM && b_major == BMJ && epilogue == E) {                                                            \
    constexpr bool kCanDyn = E != STK_EPI_BIAS_RESID_LN && E != STK_EPI_BIAS_DROP_RESID_LN;                           \
    if (kCanDyn && dyn)                                                                                               \
      return pair ? launch<AM, BMJ, E, true, kCanDyn>(ma, mb, mc, mc2, mr, p, stream)                                 \
                  : launch<AM, BMJ, E, false, kCanDyn>(ma, mb, mc, mc2, mr, p, stream);                               \
    return pair ? launch<AM, BMJ, E, true, false>(ma, mb, mc, mc2, mr, p, stream)                                     \
                : launch<AM, BMJ, E, false, false>(ma, mb, mc, mc2, mr, p, stream);                                   \
  }
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_GELU)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_GELU_SAVE)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_GELU_SAVE_GRAD)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_RESID)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_RESID_LN)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_DROP_RESID_LN)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_TANH_F32)
  STK_GEMM_CASE(0, 0, STK_EPI_F32)
  STK_GEMM_CASE(0, 0, STK_EPI_CE_STATS)
  STK_GEMM_CASE(0, 0, STK_EPI_CE_DLOGIT)
  STK_GEMM_CASE(0, 1, STK_EPI_BIAS)
  STK_GEMM_CASE(0, 1, STK_EPI_BIAS_RESID)
  STK_GEMM_CASE(0, 1, STK_EPI_DGELU)
  STK_GEMM_CASE(0, 1, STK_EPI_MUL)
  STK_GEMM_CASE(0, 1, STK_EPI_F32)
  STK_GEMM_CASE(0, 1, STK_EPI_F32_ADD)
  STK_GEMM_CASE(1, 1, STK_EPI_F32)
  STK_GEMM_CASE(1, 1, STK_EPI_F32_ADD)
#undef STK_GEMM_CASE
  set_error("stk_gemm: unsupported combination a_major=%d b_major=%d epilogue=%d", a_major, b_major, epilogue);
  return STK_ERR_UNSUPPORTED;
}

extern "C" int stk_set_gemm_dynamic(int on) {
  const int prev = gemm_dynamic() ? 1 : 0;
  g_gemm_dynamic.store(on != 0 ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

// bring-up only: copy the clock64 timeline recorded by CTA 0 of the last STK_GEMM_DEBUG launch
extern "C" __attribute__((visibility("default"))) int stk_debug_gemm_timeline(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, stk::g_gemm_timeline, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
