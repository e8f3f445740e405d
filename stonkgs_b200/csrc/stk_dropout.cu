// stk_dropout.cu — training-mode hidden dropout around the residual LayerNorms.
//
//   stk_dropout_fwd            y = drop(x)                      embeddings output (HF:110); also the backward of every
//                                                               hidden-dropout site (dropout is linear and self-adjoint)
//   stk_dropout_resid_ln_fwd   z = drop(x) + r ; y = LN(z)      BertSelfOutput / BertOutput in train() (HF:296-298, 354-356)
// One warp per 768-wide row, the lane layout of stk_embed.cu (chunk = lane + 32 i holds columns 4*chunk .. +3, i.e.
// exactly one decision word of drop_words()).  In eval() / extraction these sites are the fused GEMM epilogue
// (STK_EPI_BIAS_RESID_LN); with dropout the dense GEMM uses the plain bias epilogue and this kernel follows it.
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"
#include "stk_rng.cuh"

namespace stk {

extern std::atomic<long long> g_launches;

constexpr int kDChunks = 6;
constexpr int kDPerLane = 24;
constexpr int kDRowWarps = 8;

__device__ __forceinline__ void d_load_row(const __nv_bfloat16* row, int lane, float (&v)[kDPerLane]) {
  const uint2* p = reinterpret_cast<const uint2*>(row);
#pragma unroll
  for (int i = 0; i < kDChunks; ++i) {
    const uint2 t = __ldg(p + lane + 32 * i);
    v[4 * i] = bf16_lo(t.x); v[4 * i + 1] = bf16_hi(t.x); v[4 * i + 2] = bf16_lo(t.y); v[4 * i + 3] = bf16_hi(t.y);
  }
}
__device__ __forceinline__ void d_store_row(__nv_bfloat16* row, int lane, const float (&v)[kDPerLane]) {
  uint2* p = reinterpret_cast<uint2*>(row);
#pragma unroll
  for (int i = 0; i < kDChunks; ++i)
    p[lane + 32 * i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
}
__device__ __forceinline__ void d_apply(float (&v)[kDPerLane], int lane, uint32_t row_key, uint32_t thr, float scale) {
  const uint32_t thr4 = drop_thr4(thr);
#pragma unroll
  for (int i = 0; i < kDChunks; ++i) {
    const uint32_t chunk = static_cast<uint32_t>(lane + 32 * i);   // columns 4*chunk .. 4*chunk + 3
    uint32_t w0, w1;
    drop_words(row_key, chunk >> 1, w0, w1);
    const uint32_t signs = drop_signs((chunk & 1u) ? w1 : w0, thr4);
#pragma unroll
    for (int k = 0; k < 4; ++k) v[4 * i + k] = drop_keep(signs, k) ? v[4 * i + k] * scale : 0.f;
  }
}

__global__ void __launch_bounds__(256) dropout_fwd_kernel(const __nv_bfloat16* __restrict__ x, int M, uint32_t seed,
                                                          uint32_t site, uint32_t thr, __nv_bfloat16* __restrict__ y) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kDRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[kDPerLane];
  d_load_row(x + row * kHidden, lane, v);
  d_apply(v, lane, drop_row_key(seed, site, static_cast<uint32_t>(row)), thr, drop_scale(thr));
  d_store_row(y + row * kHidden, lane, v);
}

__global__ void __launch_bounds__(256)
dropout_resid_ln_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ resid, int M,
                        const float* __restrict__ gamma, const float* __restrict__ beta, uint32_t seed, uint32_t site,
                        uint32_t thr, __nv_bfloat16* __restrict__ z_out, __nv_bfloat16* __restrict__ y,
                        float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kDRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[kDPerLane], r[kDPerLane];
  d_load_row(x + row * kHidden, lane, v);
  d_load_row(resid + row * kHidden, lane, r);
  d_apply(v, lane, drop_row_key(seed, site, static_cast<uint32_t>(row)), thr, drop_scale(thr));
#pragma unroll
  for (int i = 0; i < kDPerLane; ++i) v[i] += r[i];
  if (z_out) {   // the backward LayerNorm reads the bf16 sum it is given: normalise that same rounded value
    d_store_row(z_out + row * kHidden, lane, v);
#pragma unroll
    for (int i = 0; i < kDPerLane; ++i) v[i] = __bfloat162float(__float2bfloat16_rn(v[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kDPerLane; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kDPerLane; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + kLnEps);
  if (mean_out && lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < kDChunks; ++i) {
    const float4 g = __ldg(g4 + lane + 32 * i), b = __ldg(b4 + lane + 32 * i);
    v[4 * i] = fmaf((v[4 * i] - mean) * rstd, g.x, b.x);
    v[4 * i + 1] = fmaf((v[4 * i + 1] - mean) * rstd, g.y, b.y);
    v[4 * i + 2] = fmaf((v[4 * i + 2] - mean) * rstd, g.z, b.z);
    v[4 * i + 3] = fmaf((v[4 * i + 3] - mean) * rstd, g.w, b.w);
  }
  d_store_row(y + row * kHidden, lane, v);
}

}  // namespace stk

using namespace stk;

extern "C" int stk_dropout_fwd(int device, void* stream, const void* x, int M, uint32_t seed, uint32_t site, uint32_t thr,
                               void* y) {
  STK_REQUIRE(x && y && M > 0 && thr < 128, "stk_dropout_fwd: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  dropout_fwd_kernel<<<(M + kDRowWarps - 1) / kDRowWarps, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), M, seed, site, thr, static_cast<__nv_bfloat16*>(y));
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_dropout_resid_ln_fwd(int device, void* stream, const void* x, const void* resid, int M,
                                        const float* gamma, const float* beta, uint32_t seed, uint32_t site, uint32_t thr,
                                        void* z_out, void* y, float* mean, float* rstd) {
  STK_REQUIRE(x && resid && y && gamma && beta && M > 0 && thr < 128, "stk_dropout_resid_ln_fwd: bad arguments");
  STK_REQUIRE((mean == nullptr) == (rstd == nullptr), "stk_dropout_resid_ln_fwd: mean/rstd must both be given or both NULL");
  STK_CHECK_CUDA(cudaSetDevice(device));
  dropout_resid_ln_kernel<<<(M + kDRowWarps - 1) / kDRowWarps, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(resid), M, gamma, beta, seed, site, thr,
      static_cast<__nv_bfloat16*>(z_out), static_cast<__nv_bfloat16*>(y), mean, rstd);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}
