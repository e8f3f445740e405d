// stk_common.cuh — shared device helpers for the STonKGs B200 kernels (sm_100a only).
//
// Thin inline-PTX wrappers for the Blackwell primitives the kernels use: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (TMEM alloc / mma / commit / ld) and UMMA descriptors.
// Everything here targets sm_100a; there is no fallback path for other architectures.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace stk {

// ---------------------------------------------------------------------------------------------
// problem constants of the path (BERT-base shape; reference stonkgs_model.py:96, BioBERT config)
// ---------------------------------------------------------------------------------------------
constexpr int kHidden = 768;
constexpr int kHeads = 12;
constexpr int kHeadDim = 64;
constexpr float kLnEps = 1e-12f;

// ---------------------------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

// Packed fp32x2 arithmetic (sm_100: one FADD2 / FMUL2 / FFMA2 instruction per pair of lanes' values):
// halves the instruction count of elementwise epilogue math.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack_f32x2(float lo, float hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2_t pack_u32x2(uint32_t lo, uint32_t hi) {
  f32x2_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(f32x2_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2_t add_f32x2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("add.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t mul_f32x2(f32x2_t a, f32x2_t b) {
  f32x2_t r;
  asm("mul.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2_t fma_f32x2(f32x2_t a, f32x2_t b, f32x2_t c) {
  f32x2_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// bf16x2 word -> two fp32 values (exact)
__device__ __forceinline__ f32x2_t bf16x2_to_f32x2(uint32_t v) {
  return pack_f32x2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
}
__device__ __forceinline__ uint32_t f32x2_to_bf16x2(f32x2_t v) {
  float lo, hi;
  unpack_f32x2(v, lo, hi);
  return pack_bf16x2(lo, hi);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Exact (erf-form) GELU of HF BERT (hidden_act="gelu"): gelu(x) = x * Phi(x).
// Phi(-|x|) = 0.5 * erfc(|x| / sqrt2) is evaluated with Abramowitz & Stegun 7.1.26
//   erfc(z) = t (a1 + t (a2 + t (a3 + t (a4 + t a5)))) exp(-z^2),  t = 1 / (1 + 0.3275911 z),  |err| < 1.5e-7
// (absolute error on Phi < 1e-7, i.e. three orders of magnitude below the bf16 resolution of the
// activations it feeds), using one MUFU.RCP + one MUFU.EX2 + 9 FMA-pipe ops, and
//   gelu(x) = max(x, 0) - |x| * Phi(-|x|)      (no branch, no select).
__device__ __forceinline__ float phi_neg_abs(float a, float& e) {  // Phi(-a) for a >= 0; e = exp(-a^2 / 2)
  const float t = fast_rcp(fmaf(0.3275911f * 0.7071067811865476f, a, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  p *= t;
  e = fast_exp2((-0.5f * 1.4426950408889634f) * a * a);
  return p * e;
}
// Forward: Phi(-a) = exp2(q(a)) with q the degree-5 minimax fit of log2(0.5 erfc(a / sqrt2)) on [0, 5] (a = min(|x|, 5);
// beyond 5, |x| Phi(-|x|) < 1.5e-6).  Relative error of Phi(-a) <= 1.6e-4 over the whole range, i.e. of gelu(x) =
// max(x, 0) - a Phi(-a) everywhere: a twelfth of the bf16 half-ulp (1.95e-3) of the activation it is stored as; absolute
// error <= 2.3e-5.  ONE MUFU.EX2 + 5 FFMA + 2 FMNMX + 1 FFMA per element.  (Round 1 used Abramowitz & Stegun 7.1.28,
// erfc = 1 / (1 + a1 z + ... + a6 z^6)^16: 6 FFMA + MUFU.RCP + 4 FMUL — five more instructions per element in an epilogue
// that is issue-bound (≈21 SASS instructions per element made FFN1 the slowest encoder GEMM), for an absolute error of
// 7e-7 that the bf16 store cannot show; its relative error on small outputs was 3e-4 as well.)
__device__ __forceinline__ float phi_neg_exp2(float a) {   // a in [0, 5]
  float q = fmaf(-0.0002699168981052935f, a, 0.0052973381243646145f);
  q = fmaf(q, a, -0.04631929472088814f);
  q = fmaf(q, a, -0.4669075310230255f);
  q = fmaf(q, a, -1.1477888822555542f);
  q = fmaf(q, a, -1.0002284049987793f);
  return fast_exp2(q);
}
__device__ __forceinline__ float gelu_erf(float x) {
  const float a = fminf(fabsf(x), 5.0f);
  return fmaf(-a, phi_neg_exp2(a), fmaxf(x, 0.0f));
}
// gelu(x) and d/dx gelu(x) together (training forward of BertIntermediate: the derivative is stored instead of the
// pre-activation): Phi(-|x|) serves both, the derivative adds one exponential for the Gaussian factor.
__device__ __forceinline__ void gelu_erf_with_grad(float x, float& y, float& dy) {
  const float ax = fabsf(x);
  const float a = fminf(ax, 5.0f);
  const float q = phi_neg_exp2(a);              // Phi(-|x|)
  y = fmaf(-a, q, fmaxf(x, 0.0f));
  const float e = fast_exp2((-0.5f * 1.4426950408889634f) * ax * ax);
  const float cdf = x >= 0.0f ? 1.0f - q : q;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);
}
// d/dx gelu(x) = Phi(x) + x * phi(x), phi(x) = exp(-x^2/2) / sqrt(2 pi): the exponential is shared with
// the erfc evaluation (2 MUFU ops per element in total)
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float a = fabsf(x);
  float e;
  const float q = phi_neg_abs(a, e);
  const float cdf = x >= 0.0f ? 1.0f - q : q;
  return fmaf(x, 0.3989422804014327f * e, cdf);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// non-blocking probe (never suspends the thread): look-ahead polls between tcgen05.mma issues
__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a pipeline bug must never hang the GPU box.  mbarrier.try_wait already suspends the
// thread for a hardware-defined interval, so the loop polls it directly; the (cheap, SM-local)
// cycle counter is only consulted every 1024 failed polls, and after ~2^33 cycles (> 4 s) the kernel
// traps (sticky launch failure reported to the host) instead of spinning forever.
// NOTE: %globaltimer must not be read on this path — it costs on the order of a microsecond.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint32_t polls = 0;
  long long t0 = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++polls & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > (1ll << 33)) {
        printf("stk: mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x, threadIdx.x,
               (void*)bar, parity);
        __trap();
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters / distributed shared memory
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// address of the same shared-memory location in CTA `rank` of this cluster (shared::cluster window)
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t cluster_addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(cluster_addr), "f"(a), "f"(b) : "memory");
}
// Asynchronous distributed-shared-memory store that also completes 8 transaction bytes on an mbarrier of
// the destination CTA: the data is visible to whoever observes that barrier's phase complete, with no
// fence on the sending side (a release.cluster arrive costs a MEMBAR.ALL.GPU, ~2000 cycles under load).
__device__ __forceinline__ void st_async_f32x2(uint32_t cluster_addr, float a, float b, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f32 [%0], {%1, %2}, [%3];" ::"r"(
                   cluster_addr),
               "f"(a), "f"(b), "r"(cluster_mbar)
               : "memory");
}
// arrive on an mbarrier of a peer CTA; release at cluster scope publishes this thread's earlier
// (and, through __syncwarp, its warp's) distributed-shared-memory stores to the waiter
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded like mbar_wait (a protocol bug must trap, not hang the box)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  uint32_t polls = 0;
  long long t0 = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if ((++polls & 1023u) == 0) {
      const long long now = clock64();
      if (t0 == 0) t0 = now;
      if (now - t0 > (1ll << 33)) {
        printf("stk: cluster mbarrier wait timeout (block %d thread %d bar %p parity %u)\n", blockIdx.x, threadIdx.x,
               (void*)bar, parity);
        __trap();
      }
    }
  }
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiles
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// global -> shared, completion signalled on an mbarrier (complete_tx::bytes)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// shared -> global (bulk group)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
// shared -> global with element-wise add (fp32): split-K partial sums and gradient accumulation
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate, one CTA.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Whole-warp forms: every lane of a converged warp executes the statement and elect.sync picks the issuing
// lane INSIDE it, so ptxas emits a plain `ELECT P; @P UTCHMMA` with uniform-register operands.  (Branching
// on an elect result cached in a register hides from ptxas that exactly one lane is active: it then wraps
// each tcgen05 instruction in an elect / issue / "any lane left?" loop and re-broadcasts its operands.)
__device__ __forceinline__ void umma_bf16_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                    uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair_warp(uint64_t* bar, uint32_t cta_mask) {
  asm volatile(
      "{\n\t.reg .pred e;\n\t.reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %1;\n\t"
      "elect.sync _|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta_mask)
      : "memory");
}
// One k-block (four K=16 steps) of a GEMM mainloop as ONE statement: a single elect, four MMAs, a
// non-blocking probe of the next slot's "full" barrier between them (returned), and the commit that frees
// this slot.  Keeps the issuing warp's instruction count per k-block minimal — it shares its scheduler with
// epilogue warps, and every cycle it waits for an issue slot beyond the 128 of the MMA in flight idles the
// tensor pipe.  PAIR selects the cta_group::2 forms (commit multicast to both CTAs of the pair).
template <bool PAIR>
__device__ __forceinline__ bool umma_kblock_warp(uint32_t d_tmem, uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1,
                                                 uint64_t a2, uint64_t b2, uint64_t a3, uint64_t b3, uint32_t idesc,
                                                 uint32_t acc_first, uint64_t* slot_free_bar, uint32_t cta_mask,
                                                 uint64_t* next_full_bar, uint32_t next_parity) {
  uint32_t ready;
  if (PAIR) {
    asm volatile(
        "{\n\t.reg .pred p, e, r;\n\t.reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %13;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], %2, %3, %10, p;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], %4, %5, %10, 1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 r, [%14], %15;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], %6, %7, %10, 1;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%1], %8, %9, %10, 1;\n\t"
        "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%12], lo;\n\t"
        "selp.u32 %0, 1, 0, r;\n\t}"
        : "=r"(ready)
        : "r"(d_tmem), "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "l"(a3), "l"(b3), "r"(idesc),
          "r"(acc_first), "r"(smem_u32(slot_free_bar)), "r"(cta_mask), "r"(smem_u32(next_full_bar)), "r"(next_parity)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p, e, r;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %11, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %2, %3, %10, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %4, %5, %10, 1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 r, [%13], %14;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %6, %7, %10, 1;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%1], %8, %9, %10, 1;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%12];\n\t"
        "selp.u32 %0, 1, 0, r;\n\t}"
        : "=r"(ready)
        : "r"(d_tmem), "l"(a0), "l"(b0), "l"(a1), "l"(b1), "l"(a2), "l"(b2), "l"(a3), "l"(b3), "r"(idesc),
          "r"(acc_first), "r"(smem_u32(slot_free_bar)), "r"(smem_u32(next_full_bar)), "r"(next_parity)
        : "memory");
  }
  return ready != 0;
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A operand read from tensor memory (lane = row, each 32-bit
// column holds two consecutive K elements), e.g. the bf16 softmax probabilities of attention.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster whose ranks differ in bit 0 run ONE 256-row MMA;
// each holds its 128 rows of A and of the accumulator, and half of B (the halves are exchanged by the
// tensor cores), which halves the shared-memory operand traffic per SM.  The even CTA issues the MMAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the pair-rank bit of a shared::cluster address
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, uint32_t ncols) {  // one warp of EACH CTA
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all earlier MMAs of this thread are complete) on the barrier at the same offset in every
// CTA of the cluster whose rank bit is set in cta_mask
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint32_t cta_mask) {
  asm volatile(
      "{\n\t.reg .b16 lo, hi;\n\t"
      "mov.b32 {lo, hi}, %1;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], lo;\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta_mask)
      : "memory");
}
// TMA load issued by either CTA of a pair; the transaction bytes complete on the barrier of the EVEN CTA
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// plain (release.cta) arrive on a barrier of another CTA of the cluster: no GPU-scope fence is emitted
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

// TMEM -> registers: thread i of the warp reads lane (base_lane + i), 32 consecutive columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> TMEM (same shape as the load above)
__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),
        "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),
        "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (sm_100 layout; bit positions as in CUTLASS cute/arch/mma_sm100_desc.hpp)
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle:
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1   [61,64) layout type (2 = SWIZZLE_128B)
// K-major operand tile  [rows][64 bf16]: 8-row groups are 1024 B apart (SBO = 1024, LBO unused).
// MN-major operand tile [k][64 bf16] x chunks: SBO = 1024 (8 k-rows), LBO = bytes between 64-wide MN chunks.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;
  return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.
//   [4,6) D format (1 = f32)  [7,10) A format (1 = bf16)  [10,13) B format (1 = bf16)
//   [15] A major (1 = MN)  [16] B major (1 = MN)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// One lane of a fully converged warp (warp-uniform control flow keeps descriptor math on the
// uniform datapath; only the issuing instruction is predicated).
// ---------------------------------------------------------------------------------------------
// Cluster launch control (sm_100): a running CTA / cluster cancels one that has not been launched yet and takes over
// its block index — hardware work stealing for persistent kernels launched with grid = number of work items.
// The 16-byte response lands in shared memory and completes 16 transaction bytes on the mbarrier.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void clc_try_cancel(uint32_t resp_smem, uint32_t bar_smem) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
               ::"r"(resp_smem), "r"(bar_smem) : "memory");
}
// the response and the mbarrier signal go to the same offsets in EVERY CTA of the cluster (issue from one CTA only)
__device__ __forceinline__ void clc_try_cancel_multicast(uint32_t resp_smem, uint32_t bar_smem) {
  asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
               ::"r"(resp_smem), "r"(bar_smem) : "memory");
}
// returns the x block index of the first CTA of the cancelled cluster, or -1 when nothing was left to cancel
__device__ __forceinline__ int clc_decode(const uint4& r) {
  const unsigned long long lo = (static_cast<unsigned long long>(r.y) << 32) | r.x;
  const unsigned long long hi = (static_cast<unsigned long long>(r.w) << 32) | r.z;
  uint32_t ok, x;
  asm volatile(
      "{\n\t.reg .b128 R;\n\t.reg .pred p;\n\t"
      "mov.b128 R, {%2, %3};\n\t"
      "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p, R;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "mov.u32 %1, 0;\n\t"
      "@p clusterlaunchcontrol.query_cancel.get_first_ctaid::x.b32.b128 %1, R;\n\t}"
      : "=r"(ok), "=r"(x)
      : "l"(lo), "l"(hi));
  return ok ? static_cast<int>(x) : -1;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// Register re-balancing between the warpgroups of a CTA (every warp of the warpgroup must execute it).
template <int N>
__device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace stk
