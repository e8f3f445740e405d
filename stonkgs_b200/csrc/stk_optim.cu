// stk_optim.cu — fused gradient-norm clip + AdamW + bf16 weight refresh.
//
// Replaces, for the live parameters of the path, what HF Trainer runs after backward
// (reference stonkgs_pretraining.py:171-193 -> transformers Trainer defaults):
//   torch.nn.utils.clip_grad_norm_(params, max_grad_norm=1.0)  and  torch.optim.AdamW.step()
// as one multi-tensor pass over (param, grad, exp_avg, exp_avg_sq): read 16 B, write 12 B per
// element, plus the 2-byte bf16 copy the GEMMs consume (so no separate cast pass is needed after the
// step).  HBM-bound: 30 B per parameter.
//
// Arithmetic follows torch.optim.AdamW (decoupled weight decay, bias-corrected):
//   p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;
//   p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with g pre-multiplied by the clip coefficient min(1, max_norm / (||g|| + 1e-6)).
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

extern std::atomic<long long> g_launches;

constexpr int kAdamChunk = 65536;  // elements per block-chunk

// Deterministic global sum of squares: block b writes its partial into slot b of a per-launch scratch row, the block
// that arrives last adds the partials in index order and accumulates into *out.  (atomicAdd of the partials would make
// the clip coefficient — and with it every updated weight — depend on arrival order: data-parallel ranks holding
// bit-identical gradients must stay bit-identical after the step.)
constexpr int kSumsqMaxBlocks = 2048;
constexpr int kSumsqSlots = 8;
__device__ float g_sumsq_part[kSumsqSlots][kSumsqMaxBlocks];
__device__ unsigned int g_sumsq_done[kSumsqSlots];

__device__ __forceinline__ void sumsq_finish(float block_sum, int slot, float* __restrict__ out) {
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    g_sumsq_part[slot][blockIdx.x] = block_sum;
    __threadfence();
    s_last = atomicAdd(&g_sumsq_done[slot], 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ float red[8];
  float t = 0.f;   // thread i sums slots i, i + 256, ... in order; then a fixed tree
  for (unsigned i = threadIdx.x; i < gridDim.x; i += 256) t += __ldcg(&g_sumsq_part[slot][i]);
  t = warp_sum(t);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) total += red[w];
    out[0] += total;
    g_sumsq_done[slot] = 0;
  }
}

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out,
                                                    int slot) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 4;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
      s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
    } else {
      for (int64_t j = i; j < n; ++j) s = fmaf(x[j], x[j], s);
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  sumsq_finish(t, slot, out);
}

// the same over a bf16 buffer scaled by `scale` (the all-reduced wire buffer of the data-parallel path: sum over ranks,
// scale = 1 / world)
__global__ void __launch_bounds__(256) sumsq_bf16_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, float scale,
                                                         float* __restrict__ out, int slot) {
  __shared__ float red[8];
  float s = 0.f;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (vec && i + 8 <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + i));
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float a = bf16_lo(w[k]), b = bf16_hi(w[k]);
        s = fmaf(a, a, s);
        s = fmaf(b, b, s);
      }
    } else {
      for (int64_t j = i; j < min(i + 8, n); ++j) { const float a = __bfloat162float(x[j]); s = fmaf(a, a, s); }
    }
  }
  s = warp_sum(s) * scale * scale;
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
  }
  sumsq_finish(t, slot, out);
}

__global__ void __launch_bounds__(256)
adamw_kernel(const StkAdamSeg* __restrict__ segs, const int32_t* __restrict__ chunk_seg,
             const int64_t* __restrict__ chunk_off, float lr, float beta1, float beta2, float eps, float wd, float bc1,
             float bc2, const float* __restrict__ sumsq, float max_norm, float grad_scale) {
  const StkAdamSeg sg = segs[chunk_seg[blockIdx.x]];
  const int64_t off = chunk_off[blockIdx.x];
  const int64_t end = min(off + static_cast<int64_t>(kAdamChunk), sg.n);
  float clip = 1.f;
  if (sumsq != nullptr) clip = fminf(1.f, max_norm / (sqrtf(__ldg(sumsq)) + 1e-6f));
  clip *= grad_scale;   // g16 gradients are a SUM over ranks: grad_scale = 1 / world turns it into the mean
  const float step = lr / bc1;
  const float inv_sqrt_bc2 = rsqrtf(bc2);
  const float decay = 1.f - lr * wd;
  float* p = static_cast<float*>(sg.p);
  const float* g = static_cast<const float*>(sg.g);
  const __nv_bfloat16* g16 = static_cast<const __nv_bfloat16*>(sg.g16);
  float* m = static_cast<float*>(sg.m);
  float* v = static_cast<float*>(sg.v);
  __nv_bfloat16* w16 = static_cast<__nv_bfloat16*>(sg.w16);
  float* pc = static_cast<float*>(sg.p32_copy);
  for (int64_t i = off + threadIdx.x * 4; i < end; i += 256 * 4) {
    if (i + 4 <= end) {
      float4 g4;
      if (g16) {
        const uint2 t = __ldg(reinterpret_cast<const uint2*>(g16 + i));
        g4 = make_float4(bf16_lo(t.x), bf16_hi(t.x), bf16_lo(t.y), bf16_hi(t.y));
      } else {
        g4 = __ldg(reinterpret_cast<const float4*>(g + i));
      }
      float4 p4 = *reinterpret_cast<float4*>(p + i);
      float4 m4 = *reinterpret_cast<float4*>(m + i);
      float4 v4 = *reinterpret_cast<float4*>(v + i);
      float gg[4] = {g4.x * clip, g4.y * clip, g4.z * clip, g4.w * clip};
      float pp[4] = {p4.x, p4.y, p4.z, p4.w}, mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        pp[k] *= decay;
        mm[k] = beta1 * mm[k] + (1.f - beta1) * gg[k];
        vv[k] = beta2 * vv[k] + (1.f - beta2) * gg[k] * gg[k];
        const float denom = sqrtf(vv[k]) * inv_sqrt_bc2 + eps;
        pp[k] -= step * (mm[k] / denom);
      }
      *reinterpret_cast<float4*>(p + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
      *reinterpret_cast<float4*>(m + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(v + i) = make_float4(vv[0], vv[1], vv[2], vv[3]);
      if (w16) *reinterpret_cast<uint2*>(w16 + i) = make_uint2(pack_bf16x2(pp[0], pp[1]), pack_bf16x2(pp[2], pp[3]));
      if (pc) *reinterpret_cast<float4*>(pc + i) = make_float4(pp[0], pp[1], pp[2], pp[3]);
    } else {
      for (int64_t j = i; j < end; ++j) {
        const float gj = (g16 ? __bfloat162float(g16[j]) : g[j]) * clip;
        float pj = p[j] * decay;
        const float mj = beta1 * m[j] + (1.f - beta1) * gj;
        const float vj = beta2 * v[j] + (1.f - beta2) * gj * gj;
        pj -= step * (mj / (sqrtf(vj) * inv_sqrt_bc2 + eps));
        p[j] = pj; m[j] = mj; v[j] = vj;
        if (w16) w16[j] = __float2bfloat16_rn(pj);
        if (pc) pc[j] = pj;
      }
    }
  }
}

}  // namespace stk

using namespace stk;

static int next_sumsq_slot() {
  static std::atomic<unsigned> n{0};
  return static_cast<int>(n.fetch_add(1, std::memory_order_relaxed) % kSumsqSlots);
}

extern "C" int stk_sumsq(int device, void* stream, const float* x, int64_t n, float* out) {
  STK_REQUIRE(x && out && n > 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0, "stk_sumsq: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms(device)) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, out, next_sumsq_slot());
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_sumsq_bf16(int device, void* stream, const void* x_bf16, int64_t n, float scale, float* out) {
  STK_REQUIRE(x_bf16 && out && n > 0, "stk_sumsq_bf16: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms(device)) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  sumsq_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), n, scale, out, next_sumsq_slot());
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_adamw_step(int device, void* stream, const StkAdamSeg* segs_dev, const int32_t* chunk_seg_dev,
                              const int64_t* chunk_off_dev, int n_chunks, float lr, float beta1, float beta2, float eps,
                              float weight_decay, float bias_correction1, float bias_correction2,
                              const float* sumsq_dev, float max_grad_norm, float grad_scale) {
  STK_REQUIRE(segs_dev && chunk_seg_dev && chunk_off_dev && n_chunks > 0, "stk_adamw_step: bad arguments");
  STK_REQUIRE(bias_correction1 > 0.f && bias_correction2 > 0.f, "stk_adamw_step: bias corrections must be positive");
  STK_CHECK_CUDA(cudaSetDevice(device));
  adamw_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(segs_dev, chunk_seg_dev, chunk_off_dev, lr, beta1,
                                                                      beta2, eps, weight_decay, bias_correction1,
                                                                      bias_correction2, sumsq_dev, max_grad_norm, grad_scale);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}
