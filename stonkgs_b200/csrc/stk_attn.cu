// stk_attn.cu — fused masked-softmax attention for the 12-head, d=64 BERT-base encoder on sm_100a.
//
// Replaces HF modeling_bert.py:115-140 (eager_attention_forward: softmax(Q K^T * 0.125 + mask) V)
// as called from BertSelfAttention.forward (HF:168-207) for both encoders of the path:
// the frozen LM backbone (S = 256, no mask; stonkgs_model.py:178) and the joint encoder
// (S = 512, additive key-padding mask; stonkgs_model.py:204-210, HF:666-672).
//
// Forward, one CTA per (128-query tile, head, batch element), 288 threads, TWO CTAs per SM
// (~100 KB shared memory, 256 TMEM columns each) so that one CTA's tensor-core phases overlap the
// other CTA's softmax phases:
//   warp 8      TMA: Q once, K_j / V_j key blocks of 128 double-buffered, straight out of the fused QKV
//               activation [B*S, 2304] (128B-swizzled boxes); single-thread tcgen05.mma issue:
//               S_j = Q K_j^T -> TMEM cols [0,128), O += P_j V_j -> TMEM cols [128,192)
//   warps 0-7   online softmax in the log2 domain, two threads per query row (64 key columns each):
//               ONE tcgen05.ld pass over S_j, block max exchanged through shared memory, P_j written as
//               bf16 into the K-major 128B-swizzled UMMA layout (first half over the dead K_j tile),
//               fp32 running sum; the running max is only advanced — and the O accumulator rescaled
//               in TMEM (tcgen05.ld/st) — when it grows by more than 8 in log2 units ("lazy
//               rescale": P stays <= 256, exact after the final 1/sum normalisation).
// No S x S tensor ever goes to HBM.  Row log-sum-exp can be saved for the backward pass.
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

extern std::atomic<long long> g_launches;

constexpr int ATT_THREADS = 288;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kRescaleThreshold = 8.0f;  // log2 units
// sQ 16K | sK 2x16K | sV 16K | sP 32K | bias 2x2K | max/sum exchange 2K | barriers
constexpr int ATT_SMEM = 16384 * 6 + 4096 + 2048 + 128;

// Persistent: grid = 2 CTAs per SM; every CTA walks work items (q-tile, head, batch) with the q-tile
// index fastest, so CTAs that run together share K/V in L2, and the loads of the next item's
// Q / K_0 / K_1 are issued while the current item's last key block is still in its softmax.
__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap map_qkv, const float* __restrict__ key_bias, int S, int num_items,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;        // [2][128 keys][128 B]
  uint8_t* sV = smem + 49152;        // [128 keys][128 B]
  uint8_t* sP = smem + 65536;        // [2 chunks of 64 keys][128 rows][128 B]
  float* sBias = reinterpret_cast<float*>(smem + 98304);   // [2 item parity][512] key bias * log2e (clamped finite)
  float* sXch = sBias + 1024;        // [2 parity][2 halves][128 rows] block-max exchange (also final sums)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sXch + 512);
  uint64_t* bar_q = bars;
  uint64_t* bar_k = bars + 1;     // [2]
  uint64_t* bar_v = bars + 3;
  uint64_t* bar_s = bars + 4;     // S_j is in TMEM
  uint64_t* bar_sread = bars + 5; // every softmax thread holds S_j in registers: TMEM columns + K buffer free
  uint64_t* bar_p = bars + 6;     // P_j is in shared memory (and O has been rescaled if needed)
  uint64_t* bar_pv = bars + 7;    // O += P_j V_j has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nblk = S >> 7;          // key blocks per item == q-tiles per (head, batch)
  constexpr uint32_t T_S = 0, T_O = 128;

  if (warp == 8) {
    if (lane == 0) {
      if ((smem_u32(smem) & 1023u) != 0) { printf("stk attn: smem base not 1024-aligned\n"); __trap(); }
      tma_prefetch_desc(&map_qkv);
      mbar_init(bar_q, 1);
      mbar_init(bar_k, 1); mbar_init(bar_k + 1, 1);
      mbar_init(bar_v, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_sread, 256);
      mbar_init(bar_p, 256);
      mbar_init(bar_pv, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);
      constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
      const uint64_t q_desc = umma_smem_desc(smem_u32(sQ), 16, 1024);
      // running use counters -> mbarrier parities
      uint32_t n_q = 0, n_k[2] = {0, 0}, n_v = 0, n_sread = 0, n_p = 0, n_pv = 0;
      uint32_t kload = 0;   // total K blocks loaded so far (buffer = kload & 1)
      auto coords = [&](int item, int& hh, int& rb, int& qq) {
        const int qt = item % nblk, rest = item / nblk;
        hh = rest % kHeads;
        rb = (rest / kHeads) * S;
        qq = qt * 128;
      };
      auto load_q = [&](int item) {
        int hh, rb, qq; coords(item, hh, rb, qq);
        mbar_arrive_expect_tx(bar_q, 16384);
        tma_load_2d(&map_qkv, bar_q, sQ, hh * 64, rb + qq);
      };
      auto load_k = [&](int item, int j) {
        int hh, rb, qq; coords(item, hh, rb, qq);
        const int bf = kload & 1;
        mbar_arrive_expect_tx(bar_k + bf, 16384);
        tma_load_2d(&map_qkv, bar_k + bf, sK + bf * 16384, 768 + hh * 64, rb + j * 128);
        ++kload;
      };
      auto load_v = [&](int item, int j) {
        int hh, rb, qq; coords(item, hh, rb, qq);
        mbar_arrive_expect_tx(bar_v, 16384);
        tma_load_2d(&map_qkv, bar_v, sV, 1536 + hh * 64, rb + j * 128);
      };
      uint32_t kuse = 0;    // total K blocks consumed by score MMAs so far
      auto issue_scores = [&]() {
        const int bf = kuse & 1;
        mbar_wait(bar_k + bf, n_k[bf] & 1);
        ++n_k[bf];
        tc_fence_after();
        const uint64_t k_desc = umma_smem_desc(smem_u32(sK + bf * 16384), 16, 1024);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + T_S, q_desc + 2 * k, k_desc + 2 * k, idesc_s, k > 0);
        umma_commit(bar_s);
        ++kuse;
      };

      int item = blockIdx.x;
      if (item < num_items) {
        load_q(item);
        load_k(item, 0);
        if (nblk > 1) load_k(item, 1);
        load_v(item, 0);
        mbar_wait(bar_q, n_q & 1); ++n_q;
        issue_scores();
      }
      for (; item < num_items; item += gridDim.x) {
        const int next = item + gridDim.x;
        for (int j = 0; j < nblk; ++j) {
          mbar_wait(bar_sread, n_sread & 1); ++n_sread;   // S_j is in registers: S columns + its K buffer are free
          tc_fence_after();
          if (j + 2 < nblk) {
            load_k(item, j + 2);
          } else if (next < num_items) {   // tail of this item: start fetching the next item's Q / K
            if (j + 1 == nblk) {           // last block: every score MMA of this item is done -> Q is free too
              load_q(next);
              load_k(next, (nblk > 1) ? 1 : 0);
            } else {                       // j + 2 == nblk: this buffer will hold the next item's K_0
              load_k(next, 0);
            }
          }
          if (j + 1 < nblk) {
            issue_scores();                // next scores run on the tensor core while softmax j is still busy
          } else if (next < num_items) {
            mbar_wait(bar_q, n_q & 1); ++n_q;
            issue_scores();                // first scores of the next item
          }
          mbar_wait(bar_p, n_p & 1); ++n_p;              // P_j is in smem, O rescaled if needed
          mbar_wait(bar_v, n_v & 1); ++n_v;
          tc_fence_after();
#pragma unroll
          for (int kb = 0; kb < 2; ++kb) {
            const uint64_t p_desc = umma_smem_desc(smem_u32(sP + kb * 16384), 16, 1024);
            const uint64_t v_desc = umma_smem_desc(smem_u32(sV + kb * 8192), 8192, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_bf16(tmem_base + T_O, p_desc + 2 * k, v_desc + k * 128, idesc_o, (j | kb | k) > 0);
          }
          umma_commit(bar_pv);
          if (j + 1 < nblk || next < num_items) {   // V is single-buffered: refill once P_j V_j has consumed it
            mbar_wait(bar_pv, n_pv & 1);
            if (j + 1 < nblk) load_v(item, j + 1); else load_v(next, 0);
          }
          ++n_pv;
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ softmax warps ================================
    const int q = warp & 3, half = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const float k1 = 0.125f * kLog2e;      // 1/sqrt(64) (HF:156) folded with log2(e)
    uint32_t n_s = 0, n_pv = 0, n_x = 0, it = 0;

    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++it) {
      const int qt = item % nblk, rest = item / nblk;
      const int h = rest % kHeads, b = rest / kHeads;
      const int row_base = b * S, q0 = qt * 128;
      // key bias of this item's batch element (double-buffered by item parity; the per-block named
      // barrier below orders these writes before any read)
      float* bias_it = sBias + (it & 1) * 512;
      for (int i = threadIdx.x; i < S; i += 256) {
        const float bz = key_bias ? __ldg(key_bias + static_cast<int64_t>(b) * S + i) * kLog2e : 0.f;
        bias_it[i] = fmaxf(bz, -3.402823466e38f);   // finfo.min * log2e overflows to -inf: keep it finite
      }
      named_bar_sync(1, 256);
      float m2 = -INFINITY;                // running (lazily advanced) row max, log2 domain
      float l0 = 0.f, l1 = 0.f;            // running sums of exp2(x2 - m2) over this thread's columns

      for (int j = 0; j < nblk; ++j) {
        mbar_wait(bar_s, n_s & 1); ++n_s;
        tc_fence_after();
        float x[64];
        {
          uint32_t r0[32], r1[32];
          tmem_ld_32x32b_x32(t_row + T_S + half * 64, r0);
          tmem_ld_32x32b_x32(t_row + T_S + half * 64 + 32, r1);
          tmem_ld_wait();
          tc_fence_before();
          mbar_arrive(bar_sread);
          const float4* bz = reinterpret_cast<const float4*>(bias_it + j * 128 + half * 64);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 b0 = bz[c], b1 = bz[8 + c];
            x[4 * c] = fmaf(__uint_as_float(r0[4 * c]), k1, b0.x);
            x[4 * c + 1] = fmaf(__uint_as_float(r0[4 * c + 1]), k1, b0.y);
            x[4 * c + 2] = fmaf(__uint_as_float(r0[4 * c + 2]), k1, b0.z);
            x[4 * c + 3] = fmaf(__uint_as_float(r0[4 * c + 3]), k1, b0.w);
            x[32 + 4 * c] = fmaf(__uint_as_float(r1[4 * c]), k1, b1.x);
            x[32 + 4 * c + 1] = fmaf(__uint_as_float(r1[4 * c + 1]), k1, b1.y);
            x[32 + 4 * c + 2] = fmaf(__uint_as_float(r1[4 * c + 2]), k1, b1.z);
            x[32 + 4 * c + 3] = fmaf(__uint_as_float(r1[4 * c + 3]), k1, b1.w);
          }
        }
        float mx[4] = {x[0], x[1], x[2], x[3]};   // four independent chains instead of one of length 64
#pragma unroll
        for (int c = 4; c < 64; c += 4) {
          mx[0] = fmaxf(mx[0], x[c]); mx[1] = fmaxf(mx[1], x[c + 1]);
          mx[2] = fmaxf(mx[2], x[c + 2]); mx[3] = fmaxf(mx[3], x[c + 3]);
        }
        float bm = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
        float* xch = sXch + (n_x & 1) * 256; ++n_x;
        xch[half * 128 + row] = bm;
        named_bar_sync(1, 256);
        bm = fmaxf(xch[row], xch[128 + row]);
        if (j > 0) {
          // P_{j-1} V_{j-1} must be complete before O is touched and before sP is overwritten
          mbar_wait(bar_pv, n_pv & 1); ++n_pv;
          tc_fence_after();
        }
        if (bm > m2 + kRescaleThreshold) {    // also true for j == 0 (m2 = -inf)
          if (j > 0) {
            const float alpha = fast_exp2(m2 - bm);
            l0 *= alpha;
            l1 *= alpha;
            uint32_t o[32];
            tmem_ld_32x32b_x32(t_row + T_O + half * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) o[c] = __float_as_uint(__uint_as_float(o[c]) * alpha);
            tmem_st_32x32b_x32(t_row + T_O + half * 32, o);
            tmem_st_wait();
          }
          m2 = bm;
        }
        uint8_t* prow = sP + half * 16384 + row * 128;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float p0 = fast_exp2(x[g * 8 + 2 * i] - m2);
            const float p1 = fast_exp2(x[g * 8 + 2 * i + 1] - m2);
            l0 += p0;
            l1 += p1;
            w[i] = pack_bf16x2(p0, p1);
          }
          *reinterpret_cast<uint4*>(prow + ((g ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the tensor core (async proxy)
        tc_fence_before();          // the O rescale is ordered before the next MMA
        mbar_arrive(bar_p);
      }

      mbar_wait(bar_pv, n_pv & 1); ++n_pv;
      tc_fence_after();
      float* xch = sXch + (n_x & 1) * 256; ++n_x;
      xch[half * 128 + row] = l0 + l1;
      named_bar_sync(1, 256);
      const float total = xch[row] + xch[128 + row];
      const float inv = 1.0f / total;
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + T_O + half * 32, r);
      tmem_ld_wait();
      uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<int64_t>(row_base + q0 + row)) * kHidden + h * 64 + half * 32);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
          w[i] = pack_bf16x2(__uint_as_float(r[g * 8 + 2 * i]) * inv, __uint_as_float(r[g * 8 + 2 * i + 1]) * inv);
        dst[g] = make_uint4(w[0], w[1], w[2], w[3]);
      }
      if (lse_out && half == 0)
        lse_out[(static_cast<int64_t>(b) * kHeads + h) * S + q0 + row] = (m2 + log2f(total)) * kLn2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace stk

using namespace stk;

extern "C" int stk_attn_fwd(int device, void* stream, const void* qkv, const float* key_bias, int B, int S, void* out,
                            float* lse) {
  STK_REQUIRE(qkv && out && B > 0, "stk_attn_fwd: bad arguments");
  STK_REQUIRE(S == 128 || S == 256 || S == 384 || S == 512, "stk_attn_fwd: S must be 128, 256, 384 or 512 (got %d)", S);
  STK_CHECK_CUDA(cudaSetDevice(device));
  CUtensorMap map;
  int rc = make_tmap_2d(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, 3 * kHidden, static_cast<uint64_t>(B) * S,
                        3 * kHidden * 2, 64, 128);
  if (rc) return rc;
  static bool configured[64] = {};
  if (!configured[device & 63]) {
    STK_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    configured[device & 63] = true;
  }
  const int num_items = (S / 128) * kHeads * B;
  const int grid = num_items < 2 * num_sms(device) ? num_items : 2 * num_sms(device);
  attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, static_cast<cudaStream_t>(stream)>>>(
      map, key_bias, S, num_items, static_cast<__nv_bfloat16*>(out), lse);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

// (backward kernel: see stk_attn_bwd.cu)
