// stk_attn_bwd.cu — fused attention backward on tcgen05 (recompute, no S x S tensor in HBM).
//
// Autograd backward of HF modeling_bert.py:115-140 (softmax(Q K^T / 8 + mask) V) for the trainable
// joint encoder (stonkgs_model.py:204-210).  Given Q, K, V (fused QKV activation), O, dO and the
// saved row log-sum-exp:
//     P  = exp(Q K^T / 8 + bias - lse)            dV = P^T dO
//     dP = dO V^T                                 dS = P o (dP - D) / 8,  D = rowsum(dO o O)
//     dQ = dS K                                   dK = dS^T Q
//
// One CTA per (128-key block j, head, batch element) loops over the query blocks i:
//   warp 8     TMA producer (K_j, V_j once; Q_i, dO_i double-buffered)
//   warp 9     tcgen05.mma issuer (warp-uniform control flow, one elected lane issues)
//   warps 0-7  two threads per query row: read S and dP from TMEM, form P and dS in registers, write
//              both as bf16 into 128B-swizzled shared tiles that serve BOTH as K-major A operand
//              (dQ = dS K) and as MN-major A operand (dV = P^T dO, dK = dS^T Q) — no transposes.
// TMEM columns: S [0,128) | dP [128,256) | dV [256,320) | dK [320,384) | dQ (two accumulators) [384,512).
// dK/dV accumulate in TMEM across the query loop and are written once as bf16 into dQKV;
// dQ_i partial products go out as fp32 TMA reduce-adds into a workspace (summed over the key
// blocks in L2) and are converted to bf16 by a small tail kernel.
#include <atomic>
#include <stdlib.h>
#include <type_traits>

#include "stk_common.cuh"
#include "stk_host.h"
#include "stk_rng.cuh"

namespace stk {

extern std::atomic<long long> g_launches;

__device__ long long g_abw_timeline[2048];   // bring-up only (STK_ATTN_DEBUG=64): clock64 stamps of CTA 0
constexpr int ABW_THREADS = 320;   // 8 compute warps + TMA warp + MMA warp
constexpr float kL2e = 1.4426950408889634f;
constexpr int ABW_QSTAGES = 3;     // Q / dO ring depth
constexpr int ABW_SMEM = 1024 + 16384 * 2 + 16384 * 2 * ABW_QSTAGES + 32768 * 3 + 1024 + 128;   // 231 552 B of the 232 448 available

// D[b,h,s] = sum_d dO * O : one warp per token row, 16-lane groups own one head
__global__ void __launch_bounds__(256) attn_bwd_prep_kernel(const __nv_bfloat16* __restrict__ o,
                                                            const __nv_bfloat16* __restrict__ dO, int B, int S,
                                                            float* __restrict__ D) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (tok >= static_cast<int64_t>(B) * S) return;
  const int b = static_cast<int>(tok / S), s = static_cast<int>(tok % S);
  const uint2* po = reinterpret_cast<const uint2*>(o + tok * kHidden);
  const uint2* pd = reinterpret_cast<const uint2*>(dO + tok * kHidden);
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const uint2 a = __ldg(po + lane + 32 * i), g = __ldg(pd + lane + 32 * i);
    float v = bf16_lo(a.x) * bf16_lo(g.x) + bf16_hi(a.x) * bf16_hi(g.x) + bf16_lo(a.y) * bf16_lo(g.y) +
              bf16_hi(a.y) * bf16_hi(g.y);
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((lane & 15) == 0) {
      const int h = (lane >> 4) + 2 * i;  // chunk (lane + 32 i) covers elements 4c..4c+3 -> head c/16
      D[(static_cast<int64_t>(b) * kHeads + h) * S + s] = v;
    }
  }
}

// fp32 dQ accumulator [B*S, 768] -> bf16 into the Q third of dQKV [B*S, 2304]
__global__ void __launch_bounds__(256) attn_bwd_dq_cast_kernel(const float* __restrict__ dq, int64_t rows,
                                                               __nv_bfloat16* __restrict__ dqkv) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;  // one thread = 8 elements
  if (i >= rows * (kHidden / 8)) return;
  const int64_t r = i / (kHidden / 8);
  const int c = static_cast<int>(i % (kHidden / 8)) * 8;
  const float4 a = __ldg(reinterpret_cast<const float4*>(dq + r * kHidden + c));
  const float4 b = __ldg(reinterpret_cast<const float4*>(dq + r * kHidden + c) + 1);
  *reinterpret_cast<uint4*>(dqkv + r * (3 * kHidden) + c) =
      make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w), pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
}

// DROP: the forward multiplied V with Pd = P o m / (1 - p_drop) (stk_attn.cu); then dV = Pd^T dO,
// dS = P o (dP o m / (1 - p_drop) - D) / 8 with the same D = rowsum(dO o O), and the masks m are regenerated here.
//
// Persistent: grid = one CTA per SM; CTA c walks the items c, c + gridDim.x, ... with item = (key block j fastest,
// head, batch element), so the CTAs that run together share Q / dO tiles in L2.  All mbarrier parities follow flat
// counters (g = query-block iteration across items, k = item), so the pipeline never drains between items:
//   * Q / dO of the next item's first iteration are loaded into the ring during the current item's last iteration,
//   * K / V of the next item are requested as soon as the current item's last dQ MMA has read K (bar_kvfree),
//   * the dK / dV read-out and store of an item overlap those loads and the next item's first score MMAs.
// A non-persistent grid (one CTA per item, first version) spent 39 % of every CTA's lifetime outside the query loop:
// ~6.9 k cycles from entry to the first scores (TMEM allocation + the TMA round trip of four tiles, with all 148 CTAs
// of a wave bursting at once), ~1.4 k for the last dV / dK MMAs and ~2.8 k for an uncoalesced dK / dV store.
template <int DBG, bool DROP>
__global__ void __launch_bounds__(ABW_THREADS)
attn_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const __grid_constant__ CUtensorMap map_do,
                const __grid_constant__ CUtensorMap map_dq, const float* __restrict__ key_bias,
                const float* __restrict__ lse, const float* __restrict__ Dws, int S, int num_items,
                __nv_bfloat16* __restrict__ dqkv, uint32_t drop_seed, uint32_t drop_site, uint32_t drop_thr) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + 16384;
  uint8_t* sQ = smem + 32768;    // [3][16 KB]  Q / dO ring of three: a TMA load under load takes 2.5-3 k cycles, most of an
  uint8_t* sdO = smem + 81920;   // [3][16 KB]  iteration, and the next scores are wanted in the middle of the current pass
  uint8_t* sP = smem + 131072;   // [2 key chunks][128 q][128 B]
  uint8_t* sdS = smem + 163840;
  uint8_t* sStage = smem + 196608;   // fp32 dQ staging: 2 x [128 rows][32 fp32] for the TMA reduce-adds
  float* sBias = reinterpret_cast<float*>(smem + 229376);   // [2 item parity][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + 256);
  uint64_t* bar_kv = bars;
  uint64_t* bar_q = bars + 1;  // [3]
  uint64_t* bar_s = bars + 4;
  uint64_t* bar_p = bars + 5;
  uint64_t* bar_dq = bars + 6;
  uint64_t* bar_qfree = bars + 7;  // [3] Q_g / dO_g buffer released (only the TMA warp waits on these)
  uint64_t* bar_dvdk = bars + 10;  // dV / dK MMAs of iteration g have completed: the P / dS tiles may be rewritten
  uint64_t* bar_kvfree = bars + 11; // the item's last dQ MMA has read K (V was last read by its last dP): K / V reusable
  uint64_t* bar_accfree = bars + 12;  // the compute warps have read the item's dV / dK accumulators out of TMEM
  uint64_t* bar_sread = bars + 13;    // every compute thread holds S_g / dP_g in registers: their TMEM columns are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq = S >> 7;           // query blocks per item == key blocks per (head, batch)
  const bool st0 = (DBG & 64) && blockIdx.x == 0 && threadIdx.x == 0;
  if (st0) g_abw_timeline[64] = clock64();   // CTA entry

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&map_qkv);
      tma_prefetch_desc(&map_do);
      tma_prefetch_desc(&map_dq);
      mbar_init(bar_kv, 1);
      for (int r = 0; r < ABW_QSTAGES; ++r) { mbar_init(bar_q + r, 1); mbar_init(bar_qfree + r, 1); }
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 256);
      mbar_init(bar_dq, 1);
      mbar_init(bar_dvdk, 1);
      mbar_init(bar_kvfree, 1);
      mbar_init(bar_accfree, 256);
      mbar_init(bar_sread, 256);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  constexpr uint32_t T_S = 0, T_DP = 128, T_DV = 256, T_DK = 320, T_DQ = 384;
  if (st0) g_abw_timeline[65] = clock64();   // prologue done

  const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
  // item -> (key block j, head h, first token row of the batch element)
  auto decode = [&](int item, int& j, int& h, int& b) {
    j = item % nq;
    const int rest = item / nq;
    h = rest % kHeads;
    b = rest / kHeads;
  };

  if (warp == 8) {
    // ================================ TMA producer ================================
    const bool leader = elect_one();
    int g = 0;
    for (int k = 0; k < my_items; ++k) {
      int j, h, b;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), j, h, b);
      const int row_base = b * S;
      for (int i = 0; i < nq; ++i, ++g) {
        const int nb = g % ABW_QSTAGES;
        // the ring slot was last read by the dV / dK MMAs of iteration g - 3 (its (g / 3 - 1)-th use)
        if (g >= ABW_QSTAGES) mbar_wait(bar_qfree + nb, ((g / ABW_QSTAGES) - 1) & 1);
        if (leader) {
          mbar_arrive_expect_tx(bar_q + nb, 32768);
          tma_load_2d(&map_qkv, bar_q + nb, sQ + nb * 16384, h * 64, row_base + i * 128);
          tma_load_2d(&map_do, bar_q + nb, sdO + nb * 16384, h * 64, row_base + i * 128);
        }
        if (i == 0) {   // after Q_0 / dO_0 (which can go out a whole iteration early): the item's K / V
          if (k > 0) mbar_wait(bar_kvfree, (k - 1) & 1);
          if (leader) {
            mbar_arrive_expect_tx(bar_kv, 32768);
            tma_load_2d(&map_qkv, bar_kv, sK, 768 + h * 64, row_base + j * 128);
            tma_load_2d(&map_qkv, bar_kv, sV, 1536 + h * 64, row_base + j * 128);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ================================ MMA issuer ================================
    // Whole-warp issue: every tcgen05 statement is executed by the converged warp and elects its issuing lane
    // itself (see stk_common.cuh: a branch on a cached elect result costs ~180 cycles per MMA).
    const bool leader = elect_one();   // timeline stamps only
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // provably warp-uniform TMEM base
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);   // S, dP
    constexpr uint32_t idesc_t = umma_idesc_bf16(128, 64, 1, 1);    // dV = P^T dO, dK = dS^T Q
    constexpr uint32_t idesc_q = umma_idesc_bf16(128, 64, 0, 1);    // dQ = dS K
    const uint64_t k_desc = umma_smem_desc(smem_u32(sK), 16, 1024);
    const uint64_t v_desc = umma_smem_desc(smem_u32(sV), 16, 1024);
    const uint64_t kT_desc = umma_smem_desc(smem_u32(sK), 8192, 1024);
    const uint64_t pT_desc = umma_smem_desc(smem_u32(sP), 16384, 1024);
    const uint64_t dsT_desc = umma_smem_desc(smem_u32(sdS), 16384, 1024);
    const uint64_t ds_desc = umma_smem_desc(smem_u32(sdS), 16, 1024);
    const uint64_t q_desc0 = umma_smem_desc(smem_u32(sQ), 16, 1024);
    const uint64_t do_desc0 = umma_smem_desc(smem_u32(sdO), 16, 1024);
    const uint64_t qT_desc0 = umma_smem_desc(smem_u32(sQ), 8192, 1024);
    const uint64_t doT_desc0 = umma_smem_desc(smem_u32(sdO), 8192, 1024);
    const int total = my_items * nq;
    int sg = 0, si = 0, sk = 0;   // next scores to issue: flat iteration, iteration within its item, item
    auto issue_scores = [&]() {   // S = Q K_j^T and dP = dO V_j^T of iteration sg into their TMEM columns
      const uint64_t boff = static_cast<uint64_t>((sg % ABW_QSTAGES) * (16384 >> 4));
      if (si == 0) mbar_wait(bar_kv, sk & 1);          // first iteration of an item: its K / V
      mbar_wait(bar_q + (sg % ABW_QSTAGES), (sg / ABW_QSTAGES) & 1);
      tc_fence_after();
      {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_u + T_S, q_desc0 + boff + 2 * k, k_desc + 2 * k, idesc_s, k > 0);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_warp(tmem_u + T_DP, do_desc0 + boff + 2 * k, v_desc + 2 * k, idesc_s, k > 0);
        umma_commit_warp(bar_s);
      }
      ++sg;
      if (++si == nq) { si = 0; ++sk; }
    };
    auto stamp = [&](int g, int slot) {
      if ((DBG & 64) && blockIdx.x == 0 && leader && g < 4) g_abw_timeline[g * 16 + slot] = clock64();
    };
    stamp(0, 0);
    if (total > 0) issue_scores();
    int i = 0, k = 0;
    for (int g = 0; g < total; ++g) {
      const uint64_t boff = static_cast<uint64_t>((g % ABW_QSTAGES) * (16384 >> 4));
      const bool last = i == nq - 1;
      stamp(g, 1);
      // Inside an item the next scores are issued as soon as the compute warps have S_g / dP_g in registers, i.e. in
      // the middle of their P / dS pass: S_{g+1} / dP_{g+1} are then ready when that pass ends.  At the end of an item
      // they follow dV / dK instead: the item's epilogue waits for those, and the next item's K / V are still in flight.
      if (!last && g + 1 < total) {
        mbar_wait(bar_sread, g & 1);
        tc_fence_after();
        issue_scores();
      }
      mbar_wait(bar_p, g & 1);           // P_g, dS_g are in smem
      tc_fence_after();
      stamp(g, 2);
      {
        // dQ_g first, into accumulator g & 1: the compute warps drain it one iteration later, behind their next P / dS
        // pass (the other accumulator was drained before P_g was handed over, so no extra barrier is needed)
#pragma unroll
        for (int kb = 0; kb < 2; ++kb)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_bf16_warp(tmem_u + T_DQ + (g & 1) * 64, ds_desc + kb * (16384 >> 4) + 2 * kk, kT_desc + kb * (8192 >> 4) + kk * 128,
                      idesc_q, (kb | kk) > 0);
        umma_commit_warp(bar_dq);
        if (last) umma_commit_warp(bar_kvfree);   // every MMA that reads this item's K / V has been issued above
      }
      {
        if (i == 0 && k > 0) {   // the previous item's accumulators must have left TMEM before they are overwritten
          mbar_wait(bar_accfree, (k - 1) & 1);
          tc_fence_after();
        }
        // MN-major operands: +2048 B (16 rows of the reduction dimension) per k step
        umma_bf16_warp(tmem_u + T_DV, pT_desc, doT_desc0 + boff, idesc_t, i > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 1; kk < 8; ++kk) umma_bf16_warp(tmem_u + T_DV, pT_desc + kk * 128, doT_desc0 + boff + kk * 128, idesc_t, 1u);
        umma_bf16_warp(tmem_u + T_DK, dsT_desc, qT_desc0 + boff, idesc_t, i > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 1; kk < 8; ++kk) umma_bf16_warp(tmem_u + T_DK, dsT_desc + kk * 128, qT_desc0 + boff + kk * 128, idesc_t, 1u);
        umma_commit_warp(bar_dvdk);              // P / dS tiles reusable
        umma_commit_warp(bar_qfree + (g % ABW_QSTAGES));   // Q_g / dO_g slot reusable once everything above has completed
      }
      if (last && g + 1 < total) issue_scores();
      stamp(g, 3);
      if (++i == nq) { i = 0; ++k; }
    }
  } else {
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float scale = 0.125f;
    int g = 0;

    // additive key bias of an item's key block in the log2 domain, clamped finite (finfo.min * log2e would overflow
    // to -inf); double-buffered by item parity: the next item's values are fetched while the current item runs
    auto stage_bias = [&](int k) {
      if (threadIdx.x < 128 && k < my_items) {
        int j, h, b;
        decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), j, h, b);
        sBias[(k & 1) * 128 + threadIdx.x] =
            key_bias ? fmaxf(__ldg(key_bias + static_cast<int64_t>(b) * S + j * 128 + threadIdx.x) * kL2e, -3.402823466e38f) : 0.f;
      }
    };
    stage_bias(0);
    named_bar_sync(1, 256);
    // row statistics (log-sum-exp, D) of the NEXT iteration are fetched one iteration ahead: their global-load latency
    // would otherwise sit between the score load and the first exponential of every iteration
    auto stat_index = [&](int k, int i) -> int64_t {
      int j, h, b;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), j, h, b);
      return (static_cast<int64_t>(b) * kHeads + h) * S + i * 128 + row;
    };
    // dQ partial of iteration gd (accumulator gd & 1): my 32 fp32 columns -> swizzled staging tile -> TMA reduce-add into
    // the fp32 workspace at (row0, col0) of that iteration.  Requires bar_dq phase gd to have been waited for.
    int pd_col = 0, pd_row = 0;
    int n_drained = 0;
    auto drain_dq = [&](int gd, int col0, int row0) {
      // every warp stages and ships its own 32 x 32 slice (one 4 KB, 1024-aligned piece of the swizzled staging tile,
      // TMA box of 32 rows): no CTA-wide barrier in the drain
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + T_DQ + (gd & 1) * 64 + half * 32, r);
      tmem_ld_wait();
      if (n_drained > 0) {  // this warp's previous reduce-add has finished reading its staging slice
        if (lane == 0) tma_wait_group_read<0>();
        __syncwarp();
      }
      uint8_t* srow = sStage + half * 16384 + row * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c)
        *reinterpret_cast<uint4*>(srow + ((c ^ (row & 7)) << 4)) = make_uint4(r[4 * c], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&map_dq, sStage + half * 16384 + q * 4096, col0 + half * 32, row0 + q * 32);
        tma_commit_group();
      }
      ++n_drained;
    };
    float nxt_lse = 0.f, nxt_D = 0.f;
    if (my_items > 0) {
      const int64_t si0 = stat_index(0, 0);
      nxt_lse = __ldg(lse + si0);
      nxt_D = __ldg(Dws + si0);
    }

    for (int k = 0; k < my_items; ++k) {
      int j, h, b;
      decode(static_cast<int>(blockIdx.x) + k * static_cast<int>(gridDim.x), j, h, b);
      const int row_base = b * S;
      const int64_t stat_base = (static_cast<int64_t>(b) * kHeads + h) * S;
      const float* bias_k = sBias + (k & 1) * 128;
      stage_bias(k + 1);   // consumed after this item's named barriers

      for (int i = 0; i < nq; ++i, ++g) {
        const float row_lse = nxt_lse, row_D = nxt_D;
        {
          const bool more_i = i + 1 < nq;
          if (more_i || k + 1 < my_items) {
            const int64_t sn = more_i ? stat_base + (i + 1) * 128 + row : stat_index(k + 1, 0);
            nxt_lse = __ldg(lse + sn);
            nxt_D = __ldg(Dws + sn);
          }
        }
        const bool st = (DBG & 64) && blockIdx.x == 0 && threadIdx.x == 0 && g < 4;
        if (st) g_abw_timeline[g * 16 + 8] = clock64();
        mbar_wait(bar_s, g & 1);
        tc_fence_after();
        if (st) g_abw_timeline[g * 16 + 9] = clock64();
        uint8_t* prow = sP + half * 16384 + row * 128;
        uint8_t* dsrow = sdS + half * 16384 + row * 128;
        const float lse2 = row_lse * kL2e;
        const float k1 = scale * kL2e;
        const float nDs = -row_D * scale;
        const uint32_t drop_key = DROP ? drop_row_key(drop_seed, drop_site, static_cast<uint32_t>(stat_base + i * 128 + row)) : 0u;
        const float dp_scale = DROP ? scale * drop_scale(drop_thr) : scale;   // dP / 8, times 1 / (1 - p_drop) with dropout
        const uint32_t thr4 = drop_thr4(drop_thr);
        // The score / dP loads of the second 32-column half are in flight while the first half is processed (a thread
        // has two warps per scheduler to hide the TMEM round trip behind, so it is software-pipelined instead).
        uint32_t rs0[32], rd0[32], rs1[32], rd1[32];
        tmem_ld_32x32b_x32(t_row + T_S + half * 64, rs0);
        tmem_ld_32x32b_x32(t_row + T_DP + half * 64, rd0);
        tmem_ld_wait();
        tmem_ld_32x32b_x32(t_row + T_S + half * 64 + 32, rs1);
        tmem_ld_32x32b_x32(t_row + T_DP + half * 64 + 32, rd1);
        // P / dS of one 32-column half: math into registers (pw / dw = four 16-byte pieces each), stores separately, so
        // that the first half's math can run BEFORE the wait for the previous iteration's dV / dK MMAs (the only readers
        // of the single-buffered P / dS tiles): in the timeline that wait cost ~0.6 k of the ~3.5 k cycles per pair.
        auto half_math = [&](auto hh_tag, const uint32_t (&rs)[32], const uint32_t (&rd)[32], uint4 (&pw)[4], uint4 (&dw)[4]) {
          constexpr int hh = decltype(hh_tag)::value;
          const float4* bz = reinterpret_cast<const float4*>(bias_k + half * 64 + hh * 32);
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const float4 ba = bz[2 * gq], bb = bz[2 * gq + 1];
            const float bias8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
            uint32_t wp[4], wd[4];
            // keep decisions of the 8 keys j*128 + half*64 + hh*32 + gq*8 .. +7 (two 4-key words); what multiplied V in
            // the forward is P o m * dscale: the mask is applied to the packed P here, dscale to dV at the final readout
            uint32_t sg2[2] = {0u, 0u};
            if (DROP) {
              uint32_t w0, w1;
              drop_words(drop_key, static_cast<uint32_t>(j * 16 + half * 8 + hh * 4 + gq), w0, w1);
              sg2[0] = drop_signs(w0, thr4);
              sg2[1] = drop_signs(w1, thr4);
            }
            auto pair = [&](auto t_tag) {
              constexpr int t = decltype(t_tag)::value;
              const int e = gq * 8 + t * 2;
              // p = exp(s/8 + bias - lse) in the log2 domain; dS = p * (dP - D) / 8
              const float p0 = fast_exp2(fmaf(__uint_as_float(rs[e]), k1, bias8[2 * t] - lse2));
              const float p1 = fast_exp2(fmaf(__uint_as_float(rs[e + 1]), k1, bias8[2 * t + 1] - lse2));
              uint32_t dp0 = rd[e], dp1 = rd[e + 1];
              uint32_t pp = pack_bf16x2(p0, p1);
              if (DROP) {   // dropped entries: dP -> 0 (their dS is -p D / 8) and P -> 0
                dp0 &= drop_mask32<(t & 1) * 2>(sg2[t >> 1]);
                dp1 &= drop_mask32<(t & 1) * 2 + 1>(sg2[t >> 1]);
                pp &= drop_mask16x2<(t & 1)>(sg2[t >> 1]);
              }
              const float d0 = p0 * fmaf(__uint_as_float(dp0), dp_scale, nDs);
              const float d1 = p1 * fmaf(__uint_as_float(dp1), dp_scale, nDs);
              wp[t] = pp;
              wd[t] = pack_bf16x2(d0, d1);
            };
            pair(std::integral_constant<int, 0>{});
            pair(std::integral_constant<int, 1>{});
            pair(std::integral_constant<int, 2>{});
            pair(std::integral_constant<int, 3>{});
            pw[gq] = make_uint4(wp[0], wp[1], wp[2], wp[3]);
            dw[gq] = make_uint4(wd[0], wd[1], wd[2], wd[3]);
          }
        };
        auto half_store = [&](int hh, const uint4 (&pw)[4], const uint4 (&dw)[4]) {
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            const int off = ((hh * 4 + gq) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + off) = pw[gq];
            *reinterpret_cast<uint4*>(dsrow + off) = dw[gq];
          }
        };
        uint4 pw0[4], dw0[4], pw1[4], dw1[4];
        half_math(std::integral_constant<int, 0>{}, rs0, rd0, pw0, dw0);
        if (g > 0) {   // dV / dK of the previous iteration have finished reading the P / dS tiles
          mbar_wait(bar_dvdk, (g - 1) & 1);
        }
        half_store(0, pw0, dw0);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(bar_sread);   // S_g / dP_g are in registers: the next scores may overwrite their TMEM columns
        half_math(std::integral_constant<int, 1>{}, rs1, rd1, pw1, dw1);
        half_store(1, pw1, dw1);
        fence_proxy_async_smem();
        tc_fence_before();
        if (st) g_abw_timeline[g * 16 + 10] = clock64();
        if ((DBG & 64) && blockIdx.x == 0 && g < 4 && lane == 0 && warp > 0) g_abw_timeline[128 + g * 8 + warp] = clock64();
        // dQ_{g-1} completed long ago; the wait must come BEFORE this thread hands over P_g: afterwards dQ_g could
        // complete too and the barrier would be a whole phase ahead of a late waiter
        if (g > 0) mbar_wait(bar_dq, (g - 1) & 1);
        mbar_arrive(bar_p);
        // the previous iteration's dQ partial leaves TMEM while the tensor core runs dQ_g and the next scores
        if (g > 0) drain_dq(g - 1, pd_col, pd_row);
        pd_col = h * 64;
        pd_row = row_base + i * 128;
        if (st) g_abw_timeline[g * 16 + 12] = clock64();
      }
      if (st0 && k == 0) g_abw_timeline[66] = clock64();   // query loop of the first item done
      // dV_j, dK_j: accumulated over all query blocks; lane = key row.  Their last MMAs are issued after
      // the dQ MMAs, so wait for the dV / dK commit of the final iteration before reading the accumulators.
      mbar_wait(bar_dvdk, (g - 1) & 1);
      tc_fence_after();
      if (st0 && k == 0) g_abw_timeline[67] = clock64();   // last dV / dK complete
      uint32_t rv[32], rk[32];
      tmem_ld_32x32b_x32(t_row + T_DV + half * 32, rv);
      tmem_ld_32x32b_x32(t_row + T_DK + half * 32, rk);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(bar_accfree);   // the next item's first dV / dK MMAs may overwrite the accumulators
      // bf16 through a staging tile so that one store instruction covers 8 rows x 64 contiguous bytes instead of
      // 32 rows x 16 bytes.  The P tile is free here (its last reader, the final dV MMA, has completed), unlike the dQ
      // staging tiles, which the last TMA reduce-add may still be reading.
      uint8_t* stg = sP + warp * 4096;   // [dV | dK][32 rows][64 B], 16-byte chunks XOR-swizzled with (row >> 1) & 3
      const float osc = DROP ? drop_scale(drop_thr) : 1.0f;   // dV = dscale * (P o m)^T dO
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t wv[4], wk[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          wv[t] = pack_bf16x2(__uint_as_float(rv[c * 8 + 2 * t]) * osc, __uint_as_float(rv[c * 8 + 2 * t + 1]) * osc);
          wk[t] = pack_bf16x2(__uint_as_float(rk[c * 8 + 2 * t]), __uint_as_float(rk[c * 8 + 2 * t + 1]));
        }
        const int off = lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4);
        *reinterpret_cast<uint4*>(stg + off) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
        *reinterpret_cast<uint4*>(stg + 2048 + off) = make_uint4(wk[0], wk[1], wk[2], wk[3]);
      }
      __syncwarp();
      __nv_bfloat16* dst0 = dqkv + static_cast<int64_t>(row_base + j * 128 + q * 32) * (3 * kHidden) + h * 64 + half * 32;
#pragma unroll
      for (int which = 0; which < 2; ++which)
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
          const int rr = ii * 8 + (lane >> 2), c = lane & 3;
          const uint4 v = *reinterpret_cast<const uint4*>(stg + which * 2048 + rr * 64 + ((c ^ ((rr >> 1) & 3)) << 4));
          *reinterpret_cast<uint4*>(dst0 + static_cast<int64_t>(rr) * (3 * kHidden) + (which == 0 ? 2 * kHidden : kHidden) + c * 8) = v;
        }
      // the next item's first P / dS pass overwrites this tile: every warp must have finished reading its own part
      named_bar_sync(1, 256);
      if (st0 && k == 0) g_abw_timeline[68] = clock64();     // dK / dV stored
    }
    if (g > 0) {   // the last iteration's dQ partial
      mbar_wait(bar_dq, (g - 1) & 1);
      drain_dq(g - 1, pd_col, pd_row);
    }
    if (lane == 0) tma_wait_group<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
  if (st0) g_abw_timeline[69] = clock64();     // CTA exit
}

}  // namespace stk

using namespace stk;

static int attn_bwd_impl(int device, void* stream_, const void* qkv, const float* key_bias, int B, int S, const void* out,
                         const void* dout, const float* lse, float* workspace, void* dqkv, bool drop, uint32_t drop_seed,
                         uint32_t drop_site, uint32_t drop_thr) {
  STK_REQUIRE(qkv && out && dout && lse && workspace && dqkv && B > 0, "stk_attn_bwd: bad arguments");
  STK_REQUIRE(S == 128 || S == 256 || S == 384 || S == 512, "stk_attn_bwd: S must be 128, 256, 384 or 512 (got %d)", S);
  STK_CHECK_CUDA(cudaSetDevice(device));
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int64_t rows = static_cast<int64_t>(B) * S;
  float* dq_acc = workspace;                    // [rows, 768]
  float* Dws = workspace + rows * kHidden;      // [B, 12, S]
  CUtensorMap map_qkv, map_do, map_dq;
  int rc = make_tmap_2d(&map_qkv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * kHidden, rows, 3 * kHidden * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&map_do, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout, kHidden, rows, kHidden * 2, 64, 128);
  if (rc) return rc;
  rc = make_tmap_2d(&map_dq, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, kHidden, rows, kHidden * 4, 32, 32);
  if (rc) return rc;
  STK_CHECK_CUDA(cudaMemsetAsync(dq_acc, 0, sizeof(float) * rows * kHidden, stream));
  attn_bwd_prep_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(out), static_cast<const __nv_bfloat16*>(dout), B, S, Dws);
  STK_CHECK_CUDA(cudaGetLastError());
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("STK_ATTN_DEBUG");
    dbg = e ? atoi(e) : 0;
  }
  auto go = [&](auto kern) -> int {
    STK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ABW_SMEM));
    const int num_items = (S / 128) * kHeads * B;
    const int grid = num_items < persistent_sms(device) ? num_items : persistent_sms(device);
    kern<<<grid, ABW_THREADS, ABW_SMEM, stream>>>(map_qkv, map_do, map_dq, key_bias, lse, Dws, S, num_items,
                                                  static_cast<__nv_bfloat16*>(dqkv), drop_seed, drop_site, drop_thr);
    return STK_OK;
  };
  if (drop) rc = go(attn_bwd_kernel<0, true>);
  else rc = (dbg == 64) ? go(attn_bwd_kernel<64, false>) : go(attn_bwd_kernel<0, false>);
  if (rc) return rc;
  STK_CHECK_CUDA(cudaGetLastError());
  attn_bwd_dq_cast_kernel<<<static_cast<unsigned>((rows * (kHidden / 8) + 255) / 256), 256, 0, stream>>>(
      dq_acc, rows, static_cast<__nv_bfloat16*>(dqkv));
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(3, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_attn_bwd(int device, void* stream, const void* qkv, const float* key_bias, int B, int S,
                            const void* out, const void* dout, const float* lse, float* workspace, void* dqkv) {
  return attn_bwd_impl(device, stream, qkv, key_bias, B, S, out, dout, lse, workspace, dqkv, false, 0, 0, 0);
}

extern "C" int stk_attn_bwd_dropout(int device, void* stream, const void* qkv, const float* key_bias, int B, int S,
                                    const void* out, const void* dout, const float* lse, float* workspace, void* dqkv,
                                    uint32_t seed, uint32_t site, uint32_t thr) {
  STK_REQUIRE(thr < 128, "stk_attn_bwd_dropout: thr must be below 128");
  return attn_bwd_impl(device, stream, qkv, key_bias, B, S, out, dout, lse, workspace, dqkv, thr > 0, seed, site, thr);
}

// bring-up only: clock64 timeline of CTA (0,0,0) of the last STK_ATTN_DEBUG=64 backward launch
extern "C" __attribute__((visibility("default"))) int stk_debug_attn_bwd_timeline(long long* host, int n) {
  return cudaMemcpyFromSymbol(host, stk::g_abw_timeline, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
