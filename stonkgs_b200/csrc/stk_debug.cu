// stk_debug.cu — bring-up microbenchmarks (not on the product path).
//
// stk_debug_mma_rate: one thread per CTA issues `iters` back-to-back tcgen05.mma (bf16, M=128, N=n, K=16,
// operands in shared memory, SWIZZLE_128B K-major, walking the four 32-byte k-slices of a 64-wide tile
// like the GEMM mainloop) with no barriers in between, and reports cycles per MMA: the tensor pipe's
// intrinsic issue rate, the denominator the GEMM timelines are read against.
#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

__global__ void __launch_bounds__(320, 1) mma_rate_kernel(int iters, int n, int same_addr, int mode, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint64_t dummy_bar[2];
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  // operands: A 128 x 64 bf16 (16 KB) at 0, B 256 x 64 bf16 (32 KB) at 16 KB — zero-filled (values do not matter)
  for (int i = threadIdx.x; i < 49152 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(dummy_bar, 1); mbar_init(dummy_bar + 1, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 1 && (mode & 256)) {
    // whole warp, converged: elect inside the MMA statement
    const uint32_t tb = __reduce_max_sync(0xffffffffu, tmem_base);
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t b0 = umma_smem_desc(smem_u32(smem + 16384), 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t k = i & 3;
      umma_bf16_warp(tb + ((i >> 5) & 1) * 256, a0 + 2 * k, b0 + 2 * k, idesc, ((mode & 8) && i % 48 == 0) ? 0u : 1u);
      if ((i & 3) == 3 && (mode & 1)) umma_commit_warp(dummy_bar);
    }
    const long long t1 = clock64();
    umma_commit_warp(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (lane_id() == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = t2 - t0; }
  } else if (threadIdx.x == 32) {
    const uint32_t idesc = umma_idesc_bf16(128, n, 0, 0);
    const uint64_t a0 = umma_smem_desc(smem_u32(smem), 16, 1024);
    const uint64_t b0 = umma_smem_desc(smem_u32(smem + 16384), 16, 1024);
    long long sink = 0;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint32_t k = same_addr ? 0 : (i & 3);
      // mode 8: a fresh accumulation (scale-D = 0) at the start of every 12 k-blocks, like a GEMM tile
      umma_bf16(tmem_base + ((i >> 5) & 1) * 256, a0 + 2 * k, b0 + 2 * k, idesc, ((mode & 8) && i % 48 == 0) ? 0u : 1u);
      if ((i & 3) == 3) {
        if (mode & 1) umma_commit(dummy_bar);                 // like the per-stage "slot free" commit
        if (mode & 2) tc_fence_after();
        if (mode & 4) { (void)mbar_try_wait(dummy_bar + 1, 1); }   // a barrier poll between k-blocks
        if (mode & 16) { const long long t = clock64(); while (clock64() - t < 250) {} }   // a 250-cycle issue gap
        if (mode & 32) { if (mbar_try_wait(dummy_bar + 1, 0)) sink += i; }               // a poll the thread depends on
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0 + (sink == 12345);      // cycles until the last MMA was accepted
    out[blockIdx.x * 2 + 1] = t2 - t0;  // cycles until all MMAs completed
  }
  else if (warp >= 2 && (mode & 192)) {
    // background load from the other warps while the MMAs run: 64 = TMEM reads of the accumulators,
    // 128 = shared-memory stores + loads (epilogue staging traffic)
    volatile uint32_t* flag = reinterpret_cast<volatile uint32_t*>(&dummy_bar[1]);
    uint32_t acc = 0;
    for (int rep = 0; rep < iters * 2; ++rep) {
      if (mode & 64) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (rep & 7) * 32, r);
        tmem_ld_wait();
        acc += r[0];
      }
      if (mode & 128) {
        uint4* q = reinterpret_cast<uint4*>(smem + 49152 - 16384) + threadIdx.x;
        *q = make_uint4(acc, rep, 0, 0);
        acc += q[128].x;
      }
    }
    if (acc == 0x12345678u) *flag = acc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace stk

extern "C" __attribute__((visibility("default"))) int stk_debug_mma_rate(int grid, int threads, int iters, int n, int same_addr, int mode,
                                                                         long long* out_dev) {
  using namespace stk;
  STK_CHECK_CUDA(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 49152 + 1024));
  mma_rate_kernel<<<grid, threads, 49152 + 1024>>>(iters, n, same_addr, mode, out_dev);
  STK_CHECK_CUDA(cudaGetLastError());
  STK_CHECK_CUDA(cudaDeviceSynchronize());
  return STK_OK;
}
