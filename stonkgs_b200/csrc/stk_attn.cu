// stk_attn.cu — fused masked-softmax attention for the 12-head, d=64 BERT-base encoder on sm_100a.
//
// Replaces HF modeling_bert.py:115-140 (eager_attention_forward: softmax(Q K^T * 0.125 + mask) V)
// as called from BertSelfAttention.forward (HF:168-207) for both encoders of the path:
// the frozen LM backbone (S = 256, no mask; stonkgs_model.py:178) and the joint encoder
// (S = 512, additive key-padding mask; stonkgs_model.py:204-210, HF:666-672).
//
// Forward, persistent CTAs over (128-query tile, head, batch element) items, 192 threads, TWO CTAs per SM
// (~100 KB shared memory, 256 TMEM columns each) so that one CTA's tensor-core phases overlap the
// other CTA's softmax phases:
//   warp 4      TMA producer: Q, then K_j / V_j key blocks of 128 through a 3-deep / 2-deep ring, straight
//               out of the fused QKV activation [B*S, 2304] (128B-swizzled boxes)
//   warp 5      tcgen05.mma issuer (warp-uniform control flow, one elected lane issues):
//               S_j = Q K_j^T -> TMEM cols [0,128) (A, B from shared memory),
//               O += P_j V_j -> TMEM cols [192,256) with the A operand P_j read FROM TENSOR MEMORY
//   warps 0-3   online softmax in the log2 domain, ONE thread per query row (all 128 key columns of the
//               block live in its registers: no cross-thread exchange, no block-wide barrier in the
//               loop): one tcgen05.ld pass over S_j, P_j packed to bf16 and written with tcgen05.st into
//               TMEM cols [128,192) (no shared-memory round trip), fp32 running sum; the running max is
//               only advanced — and the O accumulator rescaled in TMEM — when it grows by more than 8
//               in log2 units ("lazy rescale": P stays <= 256, exact after the final 1/sum).
//               The two co-resident CTAs play the role of the two ping-pong tiles of FlashAttention-4.
// No S x S tensor ever goes to HBM.  Row log-sum-exp can be saved for the backward pass.
#include <atomic>
#include <stdlib.h>
#include <type_traits>

#include "stk_common.cuh"
#include "stk_host.h"
#include "stk_rng.cuh"

namespace stk {

extern std::atomic<long long> g_launches;

// Two warpgroups: warps 0-3 softmax (one thread per query row), warp 4 TMA, warp 5 MMA, warps 6-7 only complete the
// second warpgroup so that setmaxnreg can move registers: the CTA is launched with 128 registers per thread (2 CTAs
// per SM), warpgroup 1 drops to 56 and the softmax warpgroup rises to 200.  With a uniform budget (168) the softmax
// path spilled ~20 loop variables; two CTAs' spill lines do not fit the 28 KB of L1 left beside 2 x 103 KB of shared
// memory, so every item boundary paid a chain of L2 round trips (~5 k cycles per item, measured with the timeline).
constexpr int ATT_THREADS = 256;
constexpr int ATT_REGS_SOFTMAX = 200, ATT_REGS_OTHER = 56;
__device__ long long g_attn_timeline[4096];   // bring-up only (DBG & 64): clock64 stamps of CTA 0
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// sQ 16K | sK 3x16K | sV 2x16K | bias 2x2K | max/sum exchange 2K | barriers
constexpr int ATT_KSTAGES = 3, ATT_VSTAGES = 2;
// ... | output staging 4 x 2K (one 32-row x 64-byte tile per softmax warp)
constexpr int ATT_SMEM = 16384 * 6 + 4096 + 2048 + 256 + 8192;

// Persistent: grid = 2 CTAs per SM; every CTA walks work items (q-tile, head, batch) with the q-tile
// index fastest, so CTAs that run together share K/V in L2, and the loads of the next item's
// Q / K_0 / K_1 are issued while the current item's last key block is still in its softmax.
// DBG is a bring-up bit mask (env STK_ATTN_DEBUG; 0 in production) used to attribute time to hardware
// units: 1 skip exp2 (MUFU), 2 skip the TMEM score loads, 4 skip the P stores, 8 skip fence.proxy.async,
// 16 one mbarrier arrival per warp instead of per thread, 32 skip the block-max exchange barrier.
// DROP: training-mode dropout of the attention probabilities (HF:132): the normalised probabilities are masked and
// rescaled before they multiply V, so the row sum (and the saved log-sum-exp) stay those of the full softmax; the
// keep decisions are regenerated in the backward kernel from (seed, site, row, key) — see stk_rng.cuh.
template <int DBG, bool DROP>
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const float* __restrict__ key_bias, int S, int nq, int num_items,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out, uint32_t drop_seed, uint32_t drop_site,
                uint32_t drop_thr) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;        // [3][128 keys][128 B]
  uint8_t* sV = smem + 65536;        // [2][128 keys][128 B]
  float* sBias = reinterpret_cast<float*>(smem + 98304);   // [2 item parity][512] key bias * log2e (clamped finite)
  uint32_t* sMeta = reinterpret_cast<uint32_t*>(sBias + 1024);   // [2 item parity] 2-bit state of every 32-key group
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 1024 + 512);
  uint64_t* bar_q = bars;         // Q tile landed
  uint64_t* bar_k = bars + 1;     // [3] K block landed
  uint64_t* bar_v = bars + 4;     // [2] V block landed
  uint64_t* bar_s = bars + 6;     // S_j is in TMEM
  uint64_t* bar_sread = bars + 7; // every softmax thread holds S_j in registers: the S columns are free
  uint64_t* bar_p = bars + 8;     // P_j is in tensor memory (and O has been rescaled if needed)
  uint64_t* bar_pv = bars + 9;    // O += P_j V_j has completed
  // ring "slot free" barriers with the TMA warp as their only waiter (tcgen05.commit arrives when the
  // MMAs that read the slot have completed): a producer can never fall two phases behind on these
  uint64_t* bar_kfree = bars + 10;  // [3]
  uint64_t* bar_vfree = bars + 13;  // [2]
  uint64_t* bar_qfree = bars + 15;
  uint64_t* bar_bfull = bars + 16;  // [2] the prep warp has staged the key bias of an item
  uint64_t* bar_bfree = bars + 18;  // [2] the softmax warps have finished reading it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint8_t* sOut = smem + 16384 * 6 + 4096 + 2048 + 256;   // [4 warps][32 rows][64 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = S >> 7;          // key blocks per item; nq <= nblk query tiles per (head, batch) are computed
  constexpr uint32_t T_S = 0, T_P = 128, T_O = 192;

  if (warp == 4) {
    if (lane == 0) {
      if ((smem_u32(smem) & 1023u) != 0) { printf("stk attn: smem base not 1024-aligned\n"); __trap(); }
      tma_prefetch_desc(&map_qkv);
      mbar_init(bar_q, 1);
      for (int i = 0; i < ATT_KSTAGES; ++i) { mbar_init(bar_k + i, 1); mbar_init(bar_kfree + i, 1); }
      for (int i = 0; i < ATT_VSTAGES; ++i) { mbar_init(bar_v + i, 1); mbar_init(bar_vfree + i, 1); }
      mbar_init(bar_qfree, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_sread, 128);
      mbar_init(bar_p, 128);
      mbar_init(bar_pv, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(bar_bfull + i, 1); mbar_init(bar_bfree + i, 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // this CTA's items: blockIdx.x, blockIdx.x + gridDim.x, ...; flat key-block index g = item_idx * nblk + j
  const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  const int total = my_items * nblk;

  if (warp == 4) {
    setmaxnreg_dec<ATT_REGS_OTHER>();
    // ================================ TMA producer ================================
    // K_g -> ring slot g % 3, V_g -> ring slot g % 2, Q once per item; a slot is refilled as soon as the
    // MMA warp's tcgen05.commit reports that the MMAs reading it have completed.
    const bool leader = elect_one();
    int item = blockIdx.x;
    int hh = 0, rb = 0, qq = 0;
    auto set_item = [&](int it) {
      const int qt = it % nq, rest = it / nq;
      hh = rest % kHeads;
      rb = (rest / kHeads) * S;
      qq = qt * 128;
    };
    // kq/vq: next K / V block to load; their item coordinates advance independently
    int kg = 0, k_j = 0, k_item = item, k_hh, k_rb;
    int vg = 0, v_j = 0, v_item = item, v_hh, v_rb;
    set_item(item); k_hh = v_hh = hh; k_rb = v_rb = rb;
    auto load_k = [&]() {
      if (kg >= total) return;
      const int sl = kg % ATT_KSTAGES;
      if (leader) {
        mbar_arrive_expect_tx(bar_k + sl, 16384);
        tma_load_2d(&map_qkv, bar_k + sl, sK + sl * 16384, 768 + k_hh * 64, k_rb + k_j * 128);
      }
      ++kg;
      if (++k_j == nblk) { k_j = 0; k_item += gridDim.x; if (kg < total) { set_item(k_item); k_hh = hh; k_rb = rb; } }
    };
    auto load_v = [&]() {
      if (vg >= total) return;
      const int sl = vg % ATT_VSTAGES;
      if (leader) {
        mbar_arrive_expect_tx(bar_v + sl, 16384);
        tma_load_2d(&map_qkv, bar_v + sl, sV + sl * 16384, 1536 + v_hh * 64, v_rb + v_j * 128);
      }
      ++vg;
      if (++v_j == nblk) { v_j = 0; v_item += gridDim.x; if (vg < total) { set_item(v_item); v_hh = hh; v_rb = rb; } }
    };
    auto load_q = [&](int it) {
      set_item(it);
      if (leader) {
        mbar_arrive_expect_tx(bar_q, 16384);
        tma_load_2d(&map_qkv, bar_q, sQ, hh * 64, rb + qq);
      }
    };
    // Loads are issued in consumption order (Q_item, then K_g, V_g per key block); each waits for its
    // ring slot to be released by the MMA warp's tcgen05.commit.
    int j = 0, n_item = 0;
    for (int g = 0; g < total; ++g) {
      if (g == 0) { load_q(item); ++n_item; }
      if (g >= ATT_KSTAGES) mbar_wait(bar_kfree + g % ATT_KSTAGES, ((g / ATT_KSTAGES) - 1) & 1);
      load_k();
      if (j == 0 && g > 0) {
        // first block of a later item: Q goes out as soon as the previous item's last score MMA has read the tile,
        // i.e. BEFORE the V slot wait (which follows a whole softmax pass later): a TMA load under load takes
        // 2.5-3 k cycles, about one key-block period
        mbar_wait(bar_qfree, (n_item - 1) & 1);
        load_q(item);
        ++n_item;
      }
      if (g >= ATT_VSTAGES) mbar_wait(bar_vfree + g % ATT_VSTAGES, ((g / ATT_VSTAGES) - 1) & 1);
      load_v();
      if (++j == nblk) { j = 0; item += gridDim.x; }
    }
    __syncwarp();
  } else if (warp == 5) {
    setmaxnreg_dec<ATT_REGS_OTHER>();
    // ================================ MMA issuer ================================
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    // Whole-warp issue: every tcgen05 statement is executed by the converged warp and elects its issuing lane
    // itself (see stk_common.cuh: a branch on a cached elect result costs ~180 cycles per MMA).
    const bool leader = elect_one();   // timeline stamps only
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // provably warp-uniform TMEM base
    const uint64_t q_desc = umma_smem_desc(smem_u32(sQ), 16, 1024);
    const uint64_t k_desc0 = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_desc0 = umma_smem_desc(smem_u32(sV), 8192, 1024);
    uint32_t n_q = 0;
    int dbg_cnt = 0;
    int ks = 0, kph = 0;        // K ring slot / parity of the next score block
    int sj = 0;                 // key block index within the item of the next score block
    auto issue_scores = [&]() { // S = Q K^T for the next block; the first block of an item waits for its Q tile
      if (sj == 0) { mbar_wait(bar_q, n_q & 1); ++n_q; }
      mbar_wait(bar_k + ks, kph);
      tc_fence_after();
      if ((DBG & 64) && blockIdx.x == 0 && leader && n_q < 60) g_attn_timeline[1024 + (dbg_cnt++ & 63)] = clock64();
      {
        const uint64_t k_desc = k_desc0 + static_cast<uint64_t>(ks * (16384 >> 4));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_u + T_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k > 0);
        umma_commit_warp(bar_s);
        umma_commit_warp(bar_kfree + ks);                    // K slot reusable once these MMAs have read it
        if (sj == nblk - 1) umma_commit_warp(bar_qfree);     // last score block of the item: Q reusable
      }
      if (++ks == ATT_KSTAGES) { ks = 0; kph ^= 1; }
      if (++sj == nblk) sj = 0;
    };
    if (total > 0) issue_scores();
    int vs = 0, vph = 0, pj = 0;
    auto stamp = [&](int g, int slot) {
      if ((DBG & 64) && blockIdx.x == 0 && g < 64 && leader) g_attn_timeline[g * 16 + slot] = clock64();
    };
    for (int g = 0; g < total; ++g) {
      stamp(g, 0);
      // Inside an item the next scores are issued first (they run while softmax g is still busy).  For the last key
      // block of an item the order is reversed: the next scores need the next item's Q tile, which may still be in
      // flight, and the item's epilogue waits for this P V product.
      // (when that Q tile has already landed, the normal order keeps the next item's first scores off the critical path)
      bool scores_first = g + 1 < total;
      if (scores_first && pj == nblk - 1) {   // whichever comes first: the next Q tile or P_g
        while (true) {
          if (__all_sync(0xffffffffu, mbar_test_wait(bar_q, n_q & 1))) break;
          if (__all_sync(0xffffffffu, mbar_test_wait(bar_p, g & 1))) { scores_first = false; break; }
        }
      }
      if (scores_first) {
        mbar_wait(bar_sread, g & 1);         // S_g sits in registers: the S columns are free
        tc_fence_after();
        stamp(g, 1);
        issue_scores();
      }
      stamp(g, 2);
      mbar_wait(bar_p, g & 1);               // P_g is in tensor memory, O rescaled if needed
      stamp(g, 3);
      mbar_wait(bar_v + vs, vph);
      tc_fence_after();
      stamp(g, 4);
      {
        const uint64_t v_desc = v_desc0 + static_cast<uint64_t>(vs * (16384 >> 4));
        // 8 x (K = 16 keys): A = P columns [8k, 8k+8), B = V rows [16k, 16k+16)
        umma_bf16_ts_warp(tmem_u + T_O, tmem_u + T_P, v_desc, idesc_o, pj > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 1; k < 8; ++k) umma_bf16_ts_warp(tmem_u + T_O, tmem_u + T_P + 8 * k, v_desc + k * 128, idesc_o, 1u);
        umma_commit_warp(bar_pv);
        umma_commit_warp(bar_vfree + vs);                    // V slot reusable
      }
      stamp(g, 5);
      if (!scores_first && g + 1 < total) {
        mbar_wait(bar_sread, g & 1);
        tc_fence_after();
        issue_scores();
      }
      if (DBG & 64) { mbar_wait(bar_pv, g & 1); stamp(g, 6); }
      if (++vs == ATT_VSTAGES) { vs = 0; vph ^= 1; }
      if (++pj == nblk) pj = 0;
    }
  } else if (warp == 6) {
    setmaxnreg_dec<ATT_REGS_OTHER>();
    // ================================ key-bias prep warp ================================
    // Stages the additive key bias of the item after next's batch element (x log2e, clamped finite) and the mask of
    // key blocks that hold a biased key, one item ahead of the softmax warps: the global-load latency of the bias
    // never sits between two items.
    if (key_bias) {
      int it = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
        const int b = (item / nq) / kHeads;
        if (it >= 2) mbar_wait(bar_bfree + (it & 1), ((it >> 1) - 1) & 1);
        float* bias_it = sBias + (it & 1) * 512;
        const float4* src = reinterpret_cast<const float4*>(key_bias + static_cast<int64_t>(b) * S);
        // 2 bits per 32-key group (16 groups at S = 512): 0 clean (no bias), 1 mixed, 2 masked (every key's bias is
        // below -1e30, so its probability is exactly 0 in fp32 whatever the score).  A batch element whose keys are ALL
        // masked keeps the exact path (the reference then attends uniformly).
        uint32_t meta = 0;
        bool all_masked = true;
        for (int i0 = 0; i0 < (S >> 2); i0 += 32) {     // S / 4 float4 words; word i covers keys 4i .. 4i+3
          const int i = i0 + lane;
          const float4 v = __ldg(src + i);
          const bool any_b = v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f;
          const bool all_m = v.x < -1e30f && v.y < -1e30f && v.z < -1e30f && v.w < -1e30f;
          float4 w;   // finfo.min * log2e overflows to -inf: keep it finite
          w.x = fmaxf(v.x * kLog2e, -3.402823466e38f); w.y = fmaxf(v.y * kLog2e, -3.402823466e38f);
          w.z = fmaxf(v.z * kLog2e, -3.402823466e38f); w.w = fmaxf(v.w * kLog2e, -3.402823466e38f);
          reinterpret_cast<float4*>(bias_it)[i] = w;
          const uint32_t ba = __ballot_sync(0xffffffffu, any_b), bm = __ballot_sync(0xffffffffu, all_m);
#pragma unroll
          for (int q = 0; q < 4; ++q) {                  // 8 words = one 32-key group
            const uint32_t a8 = (ba >> (8 * q)) & 0xffu, m8 = (bm >> (8 * q)) & 0xffu;
            const uint32_t state = (m8 == 0xffu) ? 2u : (a8 ? 1u : 0u);
            all_masked = all_masked && (m8 == 0xffu);
            meta |= state << (2 * ((i0 >> 3) + q));
          }
        }
        if (all_masked) meta = 0x55555555u;
        if (lane == 0) sMeta[it & 1] = meta;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_bfull + (it & 1));   // release: the stores above are visible to the waiters
      }
    }
  } else if (warp == 7) {
    setmaxnreg_dec<ATT_REGS_OTHER>();   // idle: completes warpgroup 1 for setmaxnreg
  } else {
    setmaxnreg_inc<ATT_REGS_SOFTMAX>();
    // ================================ softmax warps ================================
    const int row = warp * 32 + lane;      // warp w may touch TMEM lanes 32w .. 32w+31
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const float k1 = 0.125f * kLog2e;      // 1/sqrt(64) (HF:156) folded with log2(e)
    uint32_t n_blk = 0, it = 0;            // n_blk = flat key-block counter (mbarrier parities)

    // Item walk without per-item divisions: (q-tile, head, batch) of blockIdx.x, advanced by the decomposition of
    // gridDim.x with carries.
    int qt = static_cast<int>(blockIdx.x) % nq, h, b;
    {
      const int rest = static_cast<int>(blockIdx.x) / nq;
      h = rest % kHeads;
      b = rest / kHeads;
    }
    const int d_qt = static_cast<int>(gridDim.x) % nq;
    const int d_h = (static_cast<int>(gridDim.x) / nq) % kHeads;
    const int d_b = (static_cast<int>(gridDim.x) / nq) / kHeads;

    // The output of an item is written while the FIRST key block of the CTA's next item is in flight: its last P V
    // product completes behind that block's score load / max / exponentials instead of stalling the softmax warps.
    bool pending = false;
    float pend_inv = 0.f, pend_lse = 0.f;
    __nv_bfloat16* pend_dst = nullptr;     // warp-uniform: output row of the warp's first query row, head column 0
    float* pend_lse_ptr = nullptr;
    uint8_t* stage = sOut + warp * 2048;   // 32 rows x 64 B, 16-byte chunks XOR-swizzled with (row >> 1) & 3
    auto flush_pending = [&]() {   // requires: the pending item's last P V product has completed
      // A thread owns one output row (128 B); storing it directly would touch 32 different lines per warp instruction.
      // Each 32-column half goes through the warp's staging tile instead, so that a store instruction covers 8 rows x
      // 64 contiguous bytes.
#pragma unroll
      for (int hh2 = 0; hh2 < 2; ++hh2) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_row + T_O + hh2 * 32, o);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            w[i] = pack_bf16x2(__uint_as_float(o[c * 8 + 2 * i]) * pend_inv, __uint_as_float(o[c * 8 + 2 * i + 1]) * pend_inv);
          *reinterpret_cast<uint4*>(stage + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2), c = lane & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(stage + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(pend_dst + static_cast<int64_t>(rr) * kHidden + hh2 * 32 + c * 8) = v;
        }
        __syncwarp();
      }
      if (lse_out) *pend_lse_ptr = pend_lse;
      pending = false;
    };

    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int row_base = b * S, q0 = qt * 128;
      const float* bias_it = sBias + (it & 1) * 512;   // staged by the prep warp (double-buffered by item parity)
      // state of every 32-key group (0 clean, 1 mixed, 2 masked; warp-uniform): clean groups take the short instruction
      // stream (no bias loads / adds), masked groups produce zeros without touching the special-function unit
      uint32_t grp_state = 0;
      if (key_bias) {
        mbar_wait(bar_bfull + (it & 1), (it >> 1) & 1);
        grp_state = sMeta[it & 1];
      }
      const uint32_t drop_key = DROP ? drop_row_key(drop_seed, drop_site, static_cast<uint32_t>((b * kHeads + h) * S + q0 + row)) : 0u;
      const uint32_t drop_thr4v = drop_thr4(drop_thr);
      float m2 = -INFINITY;                // running (lazily advanced) row max, log2 domain
      float l0 = 0.f, l1 = 0.f;            // running sums of exp2(x2 - m2)

      for (int j = 0; j < nblk; ++j) {
        const bool st = (DBG & 64) && blockIdx.x == 0 && threadIdx.x == 0 && n_blk < 64;
        if (st) g_attn_timeline[n_blk * 16 + 8] = clock64();
        mbar_wait(bar_s, n_blk & 1);
        tc_fence_after();
        if (st) g_attn_timeline[n_blk * 16 + 9] = clock64();
        // the whole score row of this key block: 4 x 32 columns, one TMEM pass
        uint32_t r[4][32];
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_ld_32x32b_x32(t_row + T_S + g * 32, r[g]);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_sread);
        const float4* bz = reinterpret_cast<const float4*>(bias_it + j * 128);
        const uint32_t st4 = (grp_state >> (8 * j)) & 0xffu;   // states of this block's four groups
        float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains
        float mc[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // clean groups: max(s * k1) = k1 * max(s)
        if (st4 == 0u) {   // whole block clean: one straight instruction stream
#pragma unroll
          for (int g = 0; g < 4; g += 2)
#pragma unroll
            for (int c = 0; c < 32; c += 4) {
              mc[0] = fmaxf(mc[0], fmaxf(__uint_as_float(r[g][c]), __uint_as_float(r[g + 1][c])));
              mc[1] = fmaxf(mc[1], fmaxf(__uint_as_float(r[g][c + 1]), __uint_as_float(r[g + 1][c + 1])));
              mc[2] = fmaxf(mc[2], fmaxf(__uint_as_float(r[g][c + 2]), __uint_as_float(r[g + 1][c + 2])));
              mc[3] = fmaxf(mc[3], fmaxf(__uint_as_float(r[g][c + 3]), __uint_as_float(r[g + 1][c + 3])));
            }
        } else
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t gs = (st4 >> (2 * g)) & 3u;
          if (gs == 0u) {
#pragma unroll
            for (int c = 0; c < 32; c += 8) {
              mc[0] = fmaxf(mc[0], fmaxf(__uint_as_float(r[g][c]), __uint_as_float(r[g][c + 4])));
              mc[1] = fmaxf(mc[1], fmaxf(__uint_as_float(r[g][c + 1]), __uint_as_float(r[g][c + 5])));
              mc[2] = fmaxf(mc[2], fmaxf(__uint_as_float(r[g][c + 2]), __uint_as_float(r[g][c + 6])));
              mc[3] = fmaxf(mc[3], fmaxf(__uint_as_float(r[g][c + 3]), __uint_as_float(r[g][c + 7])));
            }
          } else if (gs == 1u) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 b0 = bz[g * 8 + c];
              mx[0] = fmaxf(mx[0], fmaf(__uint_as_float(r[g][4 * c]), k1, b0.x));
              mx[1] = fmaxf(mx[1], fmaf(__uint_as_float(r[g][4 * c + 1]), k1, b0.y));
              mx[2] = fmaxf(mx[2], fmaf(__uint_as_float(r[g][4 * c + 2]), k1, b0.z));
              mx[3] = fmaxf(mx[3], fmaf(__uint_as_float(r[g][4 * c + 3]), k1, b0.w));
            }
          }
        }
        const float bm = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])),
                               fmaxf(fmaxf(mc[0], mc[1]), fmaxf(mc[2], mc[3])) * k1);
        if (st) g_attn_timeline[n_blk * 16 + 10] = clock64();
        // Lazy rescale.  The decision is per row, but tcgen05.ld / tcgen05.st are warp-collective
        // (.sync.aligned): when ANY row of the warp must advance its max, the whole warp runs the O
        // rescale with a per-lane factor (1.0 for rows that keep their max).
        const bool advance = bm > m2 + kRescaleThreshold;   // always true for j == 0 (m2 = -inf)
        const bool rescale = j > 0 && __any_sync(0xffffffffu, advance);
        float alpha = 1.0f;
        if (rescale) {
          alpha = advance ? fast_exp2(m2 - bm) : 1.0f;
          l0 *= alpha;
          l1 *= alpha;
        }
        if (advance) m2 = bm;
        const float nm2 = -m2;
        // The exponentials only need registers, so they are formed BEFORE waiting for the previous P V product: its
        // latency (P_{j-1} hand-over + 8 MMAs) hides behind this pass instead of stalling the softmax warps.
        // Separate instruction streams per 32-key group: clean groups run FFMA / MUFU / FADD (+ half a pack) per
        // element; one predicated stream would carry the bias moves and adds through every group.
        uint32_t pk[64];               // 128 keys -> 64 packed TMEM columns of the A operand
        auto exp_group = [&](auto g_tag, auto biased_tag) {
          constexpr int g = decltype(g_tag)::value;
          constexpr bool kBiased = decltype(biased_tag)::value;
          uint32_t dw0 = 0u, dw1 = 0u;   // keep decisions of eight keys (two words), see stk_rng.cuh
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float a0 = nm2, a1 = nm2, a2 = nm2, a3 = nm2;
            if (kBiased) {
              const float4 b0 = bz[g * 8 + c];
              a0 += b0.x; a1 += b0.y; a2 += b0.z; a3 += b0.w;
            }
            const float p0 = fast_exp2(fmaf(__uint_as_float(r[g][4 * c]), k1, a0));
            const float p1 = fast_exp2(fmaf(__uint_as_float(r[g][4 * c + 1]), k1, a1));
            const float p2 = fast_exp2(fmaf(__uint_as_float(r[g][4 * c + 2]), k1, a2));
            const float p3 = fast_exp2(fmaf(__uint_as_float(r[g][4 * c + 3]), k1, a3));
            l0 += p0 + p2;
            l1 += p1 + p3;
            if (DROP) {   // the sums above are those of the full softmax; only what multiplies V is masked
              if ((c & 1) == 0) drop_words(drop_key, static_cast<uint32_t>(j * 16 + g * 4 + (c >> 1)), dw0, dw1);
              const uint32_t signs = drop_signs((c & 1) ? dw1 : dw0, drop_thr4v);
              pk[g * 16 + 2 * c] = pack_bf16x2(p0, p1) & drop_mask16x2<0>(signs);
              pk[g * 16 + 2 * c + 1] = pack_bf16x2(p2, p3) & drop_mask16x2<1>(signs);
            } else {
              pk[g * 16 + 2 * c] = pack_bf16x2(p0, p1);
              pk[g * 16 + 2 * c + 1] = pack_bf16x2(p2, p3);
            }
          }
        };
        auto exp_dispatch = [&](auto g_tag) {
          constexpr int g = decltype(g_tag)::value;
          const uint32_t gs = (st4 >> (2 * g)) & 3u;
          if (gs == 0u) exp_group(g_tag, std::false_type{});
          else if (gs == 1u) exp_group(g_tag, std::true_type{});
          else {
#pragma unroll
            for (int c = 0; c < 16; ++c) pk[g * 16 + c] = 0u;
          }
        };
        if (st4 == 0u) {   // whole block clean: one straight stream (the groups' instructions interleave freely)
          exp_group(std::integral_constant<int, 0>{}, std::false_type{});
          exp_group(std::integral_constant<int, 1>{}, std::false_type{});
          exp_group(std::integral_constant<int, 2>{}, std::false_type{});
          exp_group(std::integral_constant<int, 3>{}, std::false_type{});
        } else {
          exp_dispatch(std::integral_constant<int, 0>{});
          exp_dispatch(std::integral_constant<int, 1>{});
          exp_dispatch(std::integral_constant<int, 2>{});
          exp_dispatch(std::integral_constant<int, 3>{});
        }
        if (n_blk > 0) {
          // the previous P V product (possibly the previous item's last) must be complete before O is
          // touched and before the P columns are overwritten
          mbar_wait(bar_pv, (n_blk - 1) & 1);
          tc_fence_after();
        }
        if (j == 0 && pending) {   // warp-uniform: the previous item's O leaves TMEM before this item's first P V
          if (st) g_attn_timeline[n_blk * 16 + 12] = clock64();
          flush_pending();
          if (st) g_attn_timeline[n_blk * 16 + 13] = clock64();
        }
        if (rescale) {
#pragma unroll 1
          for (int c0 = 0; c0 < 64; c0 += 8) {   // 8 columns at a time keeps the register peak low
            uint32_t o[8];
            tmem_ld_32x32b_x8(t_row + T_O + c0, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 8; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x32b_x8(t_row + T_O + c0, o);
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) tmem_st_32x32b_x16(t_row + T_P + g * 16, *reinterpret_cast<const uint32_t(*)[16]>(pk + g * 16));
        tmem_st_wait();
        if (st) g_attn_timeline[n_blk * 16 + 11] = clock64();
        tc_fence_before();          // P store and O rescale are ordered before the next MMA
        mbar_arrive(bar_p);
        ++n_blk;
      }

      if (key_bias) mbar_arrive(bar_bfree + (it & 1));   // the prep warp may restage this bias buffer
      const float total = l0 + l1;
      pend_inv = (DROP ? drop_scale(drop_thr) : 1.0f) / total;
      pend_lse = (m2 + log2f(total)) * kLn2;
      pend_dst = out + (static_cast<int64_t>(row_base + q0 + warp * 32)) * kHidden + h * 64;
      pend_lse_ptr = lse_out + (static_cast<int64_t>(b) * kHeads + h) * S + q0 + row;
      pending = true;
      // next item of this CTA
      qt += d_qt;
      if (qt >= nq) { qt -= nq; ++h; }
      h += d_h;
      if (h >= kHeads) { h -= kHeads; ++b; }
      b += d_b;
    }
    if (pending) {
      mbar_wait(bar_pv, (n_blk - 1) & 1);
      tc_fence_after();
      flush_pending();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace stk

using namespace stk;

static int attn_fwd_impl(int device, void* stream, const void* qkv, const float* key_bias, int B, int S, void* out,
                         float* lse, bool drop, uint32_t drop_seed, uint32_t drop_site, uint32_t drop_thr, int q_rows = 0) {
  STK_REQUIRE(qkv && out && B > 0, "stk_attn_fwd: bad arguments");
  STK_REQUIRE(S == 128 || S == 256 || S == 384 || S == 512, "stk_attn_fwd: S must be 128, 256, 384 or 512 (got %d)", S);
  if (q_rows == 0) q_rows = S;
  STK_REQUIRE(q_rows > 0 && q_rows <= S && q_rows % 128 == 0, "stk_attn_fwd: q_rows must be a multiple of 128 in (0, S] (got %d)", q_rows);
  const int nq = q_rows / 128;
  STK_CHECK_CUDA(cudaSetDevice(device));
  CUtensorMap map;
  int rc = make_tmap_2d(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * kHidden, static_cast<uint64_t>(B) * S,
                        3 * kHidden * 2, 64, 128);
  if (rc) return rc;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("STK_ATTN_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  const int num_items = nq * kHeads * B;
  const int grid = num_items < 2 * persistent_sms(device) ? num_items : 2 * persistent_sms(device);
  auto go = [&](auto kern) -> int {
    STK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    kern<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(
        map, key_bias, S, nq, num_items, static_cast<__nv_bfloat16*>(out), lse, drop_seed, drop_site, drop_thr);
    return STK_OK;
  };
  // 64 = record a clock64 timeline of CTA 0 (tools/attn_dbg.py); every other value runs the production kernel
  if (drop) rc = go(attn_fwd_kernel<0, true>);
  else rc = (dbg == 64) ? go(attn_fwd_kernel<64, false>) : go(attn_fwd_kernel<0, false>);
  if (rc) return rc;
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_attn_fwd(int device, void* stream, const void* qkv, const float* key_bias, int B, int S, void* out,
                            float* lse) {
  return attn_fwd_impl(device, stream, qkv, key_bias, B, S, out, lse, false, 0, 0, 0);
}

extern "C" int stk_attn_fwd_dropout(int device, void* stream, const void* qkv, const float* key_bias, int B, int S,
                                    void* out, float* lse, uint32_t seed, uint32_t site, uint32_t thr) {
  STK_REQUIRE(thr < 128, "stk_attn_fwd_dropout: thr must be below 128");
  return attn_fwd_impl(device, stream, qkv, key_bias, B, S, out, lse, thr > 0, seed, site, thr);
}

extern "C" int stk_attn_fwd_qrows(int device, void* stream, const void* qkv, const float* key_bias, int B, int S,
                                  int q_rows, void* out, float* lse) {
  return attn_fwd_impl(device, stream, qkv, key_bias, B, S, out, lse, false, 0, 0, 0, q_rows);
}

// (backward kernel: see stk_attn_bwd.cu)

// bring-up only: copy the clock64 timeline recorded by CTA 0 of the last STK_ATTN_DEBUG&64 launch
extern "C" __attribute__((visibility("default"))) int stk_debug_attn_timeline(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, stk::g_attn_timeline, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
