// stk_gemm.cu — persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( A[M,K] * B[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// Replaces every nn.Linear of the path and its autograd backward (reference call sites:
// stonkgs_model.py:62-73,204-217; arithmetic in HF modeling_bert.py:158-160,179-181,287-298,
// 330-356,456-468,471-485).
//
// Design (one CTA per SM, 320 threads):
//   warp 0      TMA producer: A/B tiles -> 4-stage shared-memory ring (128B-swizzled boxes)
//   warp 1      TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128x256x16, cta_group::1)
//   warps 2-9   epilogue: tcgen05.ld the 128x256 fp32 accumulator (two column halves x four lane
//               quarters), fused epilogue math, swizzled staging tile, TMA store / reduce-add
// Pipelines: smem full/empty mbarriers (TMA <-> MMA), TMEM full/empty mbarriers over TWO
// accumulator stages (2 x 256 columns = all 512 TMEM columns) so the epilogue of tile i overlaps the
// MMAs of tile i+1, and a static persistent tile scheduler (grid = #SMs).
// Operand layouts: K-major or MN-major for A and B independently (UMMA descriptor major bits), so
// forward, dgrad and wgrad read the tensors where they lie — no transposed copies.
// Tails in M, N and K are handled by TMA (zero fill on load, clipping on store).
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

constexpr int BM = 128, BN = 256, BK = 64, STAGES = 4;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int B_STAGE_BYTES = BN * BK * 2;  // 32 KB
constexpr int EPI_BUF_BYTES = 128 * 128;    // [128 rows][128 B] staging tile per epilogue group
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_SMEM_BYTES = 1024 /*align slack*/ + STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 2 * EPI_BUF_BYTES + 256;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, splits, kb_total, kb_per_split;
  StkGemmEpilogue epi;
};

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

template <int A_MN, int B_MN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
            const __grid_constant__ CUtensorMap map_c, const __grid_constant__ CUtensorMap map_c2,
            const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + STAGES * A_STAGE_BYTES;
  uint8_t* smem_epi = smem_b + STAGES * B_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + 2 * EPI_BUF_BYTES);
  uint64_t* full_bar = bars;                 // [STAGES]
  uint64_t* empty_bar = bars + STAGES;       // [STAGES]
  uint64_t* tfull_bar = bars + 2 * STAGES;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    if (EPI != STK_EPI_CE_STATS) tma_prefetch_desc(&map_c);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(full_bar + i, 1);
      mbar_init(empty_bar + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar + i, 1);
      mbar_init(tempty_bar + i, 8);  // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_items = p.m_tiles * p.n_tiles * p.splits;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    // Warp-uniform control flow (all lanes walk the loop and poll the barriers, so addresses and
    // coordinates stay on the uniform datapath); one elected lane issues the TMA instructions.
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item % p.splits;
      const int tile = item / p.splits;
      const int n0 = (tile % p.n_tiles) * BN;
      const int m0 = (tile / p.n_tiles) * BM;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty_bar + stage, phase ^ 1);
        if (leader) {
          mbar_arrive_expect_tx(full_bar + stage, A_STAGE_BYTES + B_STAGE_BYTES);
          uint8_t* sa = smem_a + stage * A_STAGE_BYTES;
          uint8_t* sb = smem_b + stage * B_STAGE_BYTES;
          const int k0 = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(&map_a, full_bar + stage, sa + i * 8192, m0 + 64 * i, k0);
          } else {
            tma_load_2d(&map_a, full_bar + stage, sa, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(&map_b, full_bar + stage, sb + i * 8192, n0 + 64 * i, k0);
          } else {
            tma_load_2d(&map_b, full_bar + stage, sb, k0, n0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // Same structure: the whole warp tracks the pipeline state, one elected lane issues tcgen05.mma /
    // tcgen05.commit, so descriptor arithmetic is uniform and no per-instruction R2UR chain forms.
    const bool leader = elect_one();
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN, B_MN);
    // K-major: 8-row groups 1024 B apart, +32 B per 16-wide k step.
    // MN-major: 64-wide chunks 8192 B apart (LBO), 8 k-rows 1024 B apart (SBO), +2048 B per k step.
    const uint64_t a_desc0 = umma_smem_desc(smem_u32(smem_a), A_MN ? 8192 : 16, 1024);
    const uint64_t b_desc0 = umma_smem_desc(smem_u32(smem_b), B_MN ? 8192 : 16, 1024);
    constexpr uint32_t a_kstep = (A_MN ? 2048 : 32) >> 4;
    constexpr uint32_t b_kstep = (B_MN ? 2048 : 32) >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t as_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int split = item % p.splits;
      const int kb0 = split * p.kb_per_split;
      const int kb1 = min(kb0 + p.kb_per_split, p.kb_total);
      mbar_wait(tempty_bar + as, as_phase ^ 1);  // epilogue has drained this accumulator stage
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(full_bar + stage, phase);
        tc_fence_after();
        if (leader) {
          const uint64_t a_desc = a_desc0 + static_cast<uint64_t>((stage * A_STAGE_BYTES) >> 4);
          const uint64_t b_desc = b_desc0 + static_cast<uint64_t>((stage * B_STAGE_BYTES) >> 4);
          if (kb == kb0) umma_bf16(d_tmem, a_desc, b_desc, idesc, 0u);
          else umma_bf16(d_tmem, a_desc, b_desc, idesc, 1u);
#pragma unroll
          for (int k = 1; k < BK / 16; ++k) umma_bf16(d_tmem, a_desc + k * a_kstep, b_desc + k * b_kstep, idesc, 1u);
          umma_commit(empty_bar + stage);  // frees the smem slot once these MMAs have read it
          if (kb == kb1 - 1) umma_commit(tfull_bar + as);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (++as == 2) { as = 0; as_phase ^= 1; }
    }
  } else {
    // ============================== epilogue warps ==============================
    const int ew = warp - 2;
    const int q = warp & 3;   // TMEM lane quarter this warp may access
    const int g = ew >> 2;    // column half of the 256-wide accumulator
    const int row = q * 32 + lane;
    uint8_t* buf = smem_epi + g * EPI_BUF_BYTES;
    const bool store_thread = ((ew & 3) == 0) && lane == 0;
    const uint32_t bar_id = 1 + g;
    const StkGemmEpilogue& e = p.epi;
    float ce_scale = 0.f;
    if (EPI == STK_EPI_CE_DLOGIT) ce_scale = __ldg(e.scale_dev);

    int as = 0;
    uint32_t as_phase = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      const int tile = item / p.splits;
      const int n_blk = tile % p.n_tiles;
      const int n0 = n_blk * BN;
      const int m0 = (tile / p.n_tiles) * BM;
      const int m = m0 + row;
      const bool m_ok = m < p.M;
      // Residual / saved pre-activation rows do not depend on the accumulator: fetch both 64-column
      // chunks of this thread's row now so the global-load latency hides behind the MMA of this tile.
      constexpr bool kPrefetchExtra = EPI == STK_EPI_BIAS_RESID || EPI == STK_EPI_DGELU;
      // COALESCED: thread t of the 128-thread group fetches 16-byte piece (t + 128 i) of the
      // [128 rows x 128 B] residual tile (8 consecutive threads = one full 128-byte row segment); the
      // pieces are transposed to the thread-per-row accumulator layout through the staging tile later.
      uint4 ex_pre[2][8];
      const int gt = (ew & 3) * 32 + lane;   // thread index within the epilogue group
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
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN + g * 128;

      // per-row CE state
      float ce_max = -INFINITY, ce_sum = 0.f;
      int label = -1;
      float row_lse = 0.f;
      if (EPI == STK_EPI_CE_STATS || EPI == STK_EPI_CE_DLOGIT) {
        if (m_ok) label = __ldg(e.labels + m) - e.n_offset;  // column within this call's B block
        if (EPI == STK_EPI_CE_DLOGIT && m_ok) row_lse = __ldg(e.lse + m);
      }

#pragma unroll
      for (int chunk = 0; chunk < 2; ++chunk) {  // 64 accumulator columns per step
        uint32_t r[2][32];
        tmem_ld_32x32b_x32(t_row + chunk * 64, r[0]);
        tmem_ld_32x32b_x32(t_row + chunk * 64 + 32, r[1]);
        tmem_ld_wait();
        if (chunk == 1) {  // accumulator fully read: hand the TMEM stage back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar + as);
        }
        const int nc = n0 + g * 128 + chunk * 64;  // first global column of this chunk

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
              if (n == label) e.tgt_logit[m] = v;
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
                if (EPI == STK_EPI_BIAS_TANH_F32) {
                  const int n = nc + h * 32 + c * 4 + i;
                  const float b = (e.bias != nullptr && n < p.N) ? __ldg(e.bias + n) : 0.f;
                  v[i] = tanhf(v[i] + b);
                }
              }
              data[c] = make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]),
                                   __float_as_uint(v[3]));
            }
            stage_and_store<EPI == STK_EPI_F32_ADD>(&map_c, buf, row, data, nc + h * 32, m0, store_thread, bar_id);
          }
          continue;
        }

        // bf16 outputs
        constexpr bool kHasBias = EPI == STK_EPI_BIAS || EPI == STK_EPI_BIAS_GELU ||
                                  EPI == STK_EPI_BIAS_GELU_SAVE || EPI == STK_EPI_BIAS_RESID;
        constexpr bool kHasExtra = EPI == STK_EPI_BIAS_RESID || EPI == STK_EPI_DGELU;
        const bool bias_vec = kHasBias && e.bias != nullptr && nc + 64 <= p.N;
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
        uint4 data[8];
        uint4 data2[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t w[4], w2[4];
          float bias_v[8];
          if (kHasBias) {
            if (bias_vec) {
              const float4 b0 = __ldg(reinterpret_cast<const float4*>(e.bias + nc) + 2 * c);
              const float4 b1 = __ldg(reinterpret_cast<const float4*>(e.bias + nc) + 2 * c + 1);
              bias_v[0] = b0.x; bias_v[1] = b0.y; bias_v[2] = b0.z; bias_v[3] = b0.w;
              bias_v[4] = b1.x; bias_v[5] = b1.y; bias_v[6] = b1.z; bias_v[7] = b1.w;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                bias_v[i] = (e.bias != nullptr && nc + c * 8 + i < p.N) ? __ldg(e.bias + nc + c * 8 + i) : 0.f;
            }
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
              v0 = m_ok ? p0 * ce_scale : 0.f;
              v1 = m_ok ? p1 * ce_scale : 0.f;
            }
            w[i] = pack_bf16x2(v0, v1);
          }
          data[c] = make_uint4(w[0], w[1], w[2], w[3]);
          if (EPI == STK_EPI_BIAS_GELU_SAVE) data2[c] = make_uint4(w2[0], w2[1], w2[2], w2[3]);
        }
        stage_and_store<false, kHasExtra>(&map_c, buf, row, data, nc, m0, store_thread, bar_id);
        if (EPI == STK_EPI_BIAS_GELU_SAVE)
          stage_and_store<false>(&map_c2, buf, row, data2, nc, m0, store_thread, bar_id);
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
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
extern std::atomic<long long> g_launches;

template <int A_MN, int B_MN, int EPI>
static int launch(const CUtensorMap& ma, const CUtensorMap& mb, const CUtensorMap& mc, const CUtensorMap& mc2,
                  const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_kernel<A_MN, B_MN, EPI>;
  static bool configured[64] = {};
  int dev = 0;
  STK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    STK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    configured[dev & 63] = true;
  }
  kern<<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, stream>>>(ma, mb, mc, mc2, p);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
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
  STK_CHECK_CUDA(cudaSetDevice(device));
  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.m_tiles = (M + BM - 1) / BM;
  p.n_tiles = (N + BN - 1) / BN;
  p.kb_total = (K + BK - 1) / BK;
  int splits = split_k < 1 ? 1 : split_k;
  if (splits > p.kb_total) splits = p.kb_total;
  if (epilogue != STK_EPI_F32_ADD) splits = 1;
  p.kb_per_split = (p.kb_total + splits - 1) / splits;
  p.splits = (p.kb_total + p.kb_per_split - 1) / p.kb_per_split;
  if (epi) p.epi = *epi;

  const bool f32_out = epilogue == STK_EPI_F32 || epilogue == STK_EPI_F32_ADD || epilogue == STK_EPI_BIAS_TANH_F32;
  const bool has_c = epilogue != STK_EPI_CE_STATS;
  if (epilogue == STK_EPI_BIAS_RESID || epilogue == STK_EPI_DGELU) {
    STK_REQUIRE(epi && epi->resid && epi->ldr % 8 == 0 && N % 64 == 0, "stk_gemm: residual epilogue needs resid, ldr%%8==0, N%%64==0");
  }
  if (epilogue == STK_EPI_BIAS_GELU_SAVE) STK_REQUIRE(epi && epi->c2 && epi->ldc2 % 8 == 0, "stk_gemm: GELU_SAVE needs c2");
  if (epilogue == STK_EPI_CE_STATS)
    STK_REQUIRE(epi && epi->labels && epi->ce_partial && epi->tgt_logit && epi->n_offset % 256 == 0, "stk_gemm: CE_STATS args");
  if (epilogue == STK_EPI_CE_DLOGIT)
    STK_REQUIRE(epi && epi->labels && epi->lse && epi->scale_dev && epi->n_offset % 256 == 0, "stk_gemm: CE_DLOGIT args");
  if (has_c) {
    STK_REQUIRE(C != nullptr && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "stk_gemm: C null or misaligned");
    STK_REQUIRE(ldc % (f32_out ? 4 : 8) == 0, "stk_gemm: ldc must be a multiple of 16 bytes");
  }

  CUtensorMap ma, mb, mc, mc2;
  int rc;
  // A: K-major -> stored [M][K], box {64 k, 128 m};  MN-major -> stored [K][M], box {64 m, 64 k}
  if (a_major == 0) rc = make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda * 2, 64, BM);
  else rc = make_tmap_2d(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda * 2, 64, 64);
  if (rc) return rc;
  if (b_major == 0) rc = make_tmap_2d(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, ldb * 2, 64, BN);
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
  if (epilogue == STK_EPI_BIAS_GELU_SAVE) {
    rc = make_tmap_2d(&mc2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, epi->c2, N, M, epi->ldc2 * 2, 64, 128);
    if (rc) return rc;
  }
  const int items = p.m_tiles * p.n_tiles * p.splits;
  const int sms = num_sms(device);
  const int grid = items < sms ? items : sms;

#define STK_GEMM_CASE(AM, BMJ, E)                                   \
  if (a_major == AM && b_major == BMJ && epilogue == E)             \
    return launch<AM, BMJ, E>(ma, mb, mc, mc2, p, grid, stream);
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_GELU)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_GELU_SAVE)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_RESID)
  STK_GEMM_CASE(0, 0, STK_EPI_BIAS_TANH_F32)
  STK_GEMM_CASE(0, 0, STK_EPI_F32)
  STK_GEMM_CASE(0, 0, STK_EPI_CE_STATS)
  STK_GEMM_CASE(0, 0, STK_EPI_CE_DLOGIT)
  STK_GEMM_CASE(0, 1, STK_EPI_BIAS)
  STK_GEMM_CASE(0, 1, STK_EPI_BIAS_RESID)
  STK_GEMM_CASE(0, 1, STK_EPI_DGELU)
  STK_GEMM_CASE(0, 1, STK_EPI_F32)
  STK_GEMM_CASE(0, 1, STK_EPI_F32_ADD)
  STK_GEMM_CASE(1, 1, STK_EPI_F32)
  STK_GEMM_CASE(1, 1, STK_EPI_F32_ADD)
#undef STK_GEMM_CASE
  set_error("stk_gemm: unsupported combination a_major=%d b_major=%d epilogue=%d", a_major, b_major, epilogue);
  return STK_ERR_UNSUPPORTED;
}
