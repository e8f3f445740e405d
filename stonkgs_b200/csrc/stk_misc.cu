// stk_misc.cu — small HBM-bound helpers around the GEMM / attention kernels: mask -> additive key
// bias, fp32 -> bf16 weight casts, labelled-row gather / scatter-add for the heads, column sums
// (bias gradients), cross-entropy finalisation and the NSP head.
//
// Reference lines: HF modeling_bert.py:666-672 (extended attention mask), stonkgs_model.py:229-245
// (three mean cross-entropies), HF:528-533 (seq_relationship head).
#include <atomic>
#include <float.h>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

extern std::atomic<long long> g_launches;

__global__ void mask_to_bias_kernel(const int64_t* __restrict__ mask, int64_t n, float* __restrict__ bias) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // (1 - mask) * finfo(float32).min, exactly as HF builds the additive mask
  if (i < n) bias[i] = (1.0f - static_cast<float>(mask[i])) * (-FLT_MAX);
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(src + i));
      const float4 b = __ldg(reinterpret_cast<const float4*>(src + i) + 1);
      *reinterpret_cast<uint4*>(dst + i) = make_uint4(pack_bf16x2(a.x, a.y), pack_bf16x2(a.z, a.w),
                                                      pack_bf16x2(b.x, b.y), pack_bf16x2(b.z, b.w));
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = __float2bfloat16_rn(src[j]);
    }
  }
}

// one warp per row of 768 bf16 (1536 B = 3 x 16 B per lane)
__global__ void __launch_bounds__(256) gather_rows_kernel(const __nv_bfloat16* __restrict__ src,
                                                          const int32_t* __restrict__ idx, int n_rows,
                                                          __nv_bfloat16* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int32_t r = __ldg(idx + row);
  uint4* d = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(row) * kHidden);
  if (r < 0) {   // padding entry of a fixed-capacity row list: zeros
#pragma unroll
    for (int i = 0; i < 3; ++i) d[lane + 32 * i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(r) * kHidden);
#pragma unroll
  for (int i = 0; i < 3; ++i) d[lane + 32 * i] = __ldg(s + lane + 32 * i);
}

__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const __nv_bfloat16* __restrict__ src,
                                                               const int32_t* __restrict__ idx, int n_rows,
                                                               __nv_bfloat16* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int32_t r = __ldg(idx + row);
  if (r < 0) return;   // padding entry
  const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<int64_t>(row) * kHidden);
  uint4* d = reinterpret_cast<uint4*>(dst + static_cast<int64_t>(r) * kHidden);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint4 a = __ldg(s + lane + 32 * i);
    uint4 b = d[lane + 32 * i];
    b.x = pack_bf16x2(bf16_lo(a.x) + bf16_lo(b.x), bf16_hi(a.x) + bf16_hi(b.x));
    b.y = pack_bf16x2(bf16_lo(a.y) + bf16_lo(b.y), bf16_hi(a.y) + bf16_hi(b.y));
    b.z = pack_bf16x2(bf16_lo(a.z) + bf16_lo(b.z), bf16_hi(a.z) + bf16_hi(b.z));
    b.w = pack_bf16x2(bf16_lo(a.w) + bf16_lo(b.w), bf16_hi(a.w) + bf16_hi(b.w));
    d[lane + 32 * i] = b;
  }
}

// column sums of a bf16 matrix: block = 8 warps over a 64-column strip and a slab of rows;
// lane owns 2 adjacent columns (128 B per warp per row), warps stride over rows.
__global__ void __launch_bounds__(256) colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t ld, int M, int N,
                                                     int rows_per_block, float* __restrict__ out) {
  __shared__ float red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * 64 + lane * 2;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(r0 + rows_per_block, M);
  float a0 = 0.f, a1 = 0.f;
  if (n < N) {
    const __nv_bfloat16* p = x + n;
    int r = r0 + warp;
    for (; r + 24 < r1; r += 32) {  // 4 independent loads in flight per lane
      uint32_t v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(reinterpret_cast<const uint32_t*>(p + static_cast<int64_t>(r + 8 * u) * ld));
#pragma unroll
      for (int u = 0; u < 4; ++u) { a0 += bf16_lo(v[u]); a1 += bf16_hi(v[u]); }
    }
    for (; r < r1; r += 8) {
      const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p + static_cast<int64_t>(r) * ld));
      a0 += bf16_lo(v); a1 += bf16_hi(v);
    }
  }
  red[warp][lane * 2] = a0;
  red[warp][lane * 2 + 1] = a1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < N) atomicAdd(out + c, s);
  }
}

// one warp per row: combine the (max, sumexp) slab partials into lse and the row loss
__global__ void __launch_bounds__(256) ce_finalize_kernel(const float2* __restrict__ part, int64_t pitch,
                                                          const float* __restrict__ tgt, int M,
                                                          float* __restrict__ lse_out, float* __restrict__ row_loss) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const float2* p = part + static_cast<int64_t>(row) * pitch;
  float mx = -INFINITY;
  for (int64_t i = lane; i < pitch; i += 32) mx = fmaxf(mx, __ldg(&p[i].x));
  mx = warp_max(mx);
  float s = 0.f;
  for (int64_t i = lane; i < pitch; i += 32) {
    const float2 v = __ldg(p + i);
    if (v.x > -INFINITY) s += v.y * exp2f((v.x - mx) * 1.4426950408889634f);
  }
  s = warp_sum(s);
  if (lane == 0) {
    const float lse = mx + logf(s);
    lse_out[row] = lse;
    if (row_loss) row_loss[row] = lse - __ldg(tgt + row);
  }
}

// one warp per sample: logits = pooled . w^T + b ; row loss = logsumexp - logit[label]
__global__ void __launch_bounds__(256) nsp_head_fwd_kernel(const float* __restrict__ pooled, int B,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           const int64_t* __restrict__ labels,
                                                           float* __restrict__ logits, float* __restrict__ row_loss,
                                                           int* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float s0 = 0.f, s1 = 0.f;
  for (int c = lane; c < kHidden; c += 32) {
    const float x = __ldg(pooled + static_cast<int64_t>(b) * kHidden + c);
    s0 = fmaf(x, __ldg(w + c), s0);
    s1 = fmaf(x, __ldg(w + kHidden + c), s1);
  }
  s0 = warp_sum(s0) + __ldg(bias);
  s1 = warp_sum(s1) + __ldg(bias + 1);
  if (lane == 0) {
    logits[2 * b] = s0;
    logits[2 * b + 1] = s1;
    if (labels && row_loss) {
      const float m = fmaxf(s0, s1);
      const float lse = m + logf(expf(s0 - m) + expf(s1 - m));
      const int64_t l = __ldg(labels + b);
      // a target outside {0, 1} raises in torch's cross-entropy: flag it; NaN keeps the loss loud and the backward
      // (which gives such a row no one-hot) consistent with what was reported
      if (l != 0 && l != 1) {
        row_loss[b] = __int_as_float(0x7fc00000);
        if (err_flag) atomicOr(err_flag, 2);
      } else {
        row_loss[b] = lse - (l == 0 ? s0 : s1);
      }
    }
  }
}


// dx = dy * gelu_erf'(pre) over n bf16 elements (head transform backward; the encoder's FFN uses the
// fused DGELU GEMM epilogue instead)
__global__ void __launch_bounds__(256) gelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                       const __nv_bfloat16* __restrict__ pre, int64_t n,
                                                       __nv_bfloat16* __restrict__ dx) {
  const int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i >= n) return;
  const uint4 a = __ldg(reinterpret_cast<const uint4*>(dy + i));
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(pre + i));
  uint4 o;
  o.x = pack_bf16x2(bf16_lo(a.x) * gelu_erf_grad(bf16_lo(u.x)), bf16_hi(a.x) * gelu_erf_grad(bf16_hi(u.x)));
  o.y = pack_bf16x2(bf16_lo(a.y) * gelu_erf_grad(bf16_lo(u.y)), bf16_hi(a.y) * gelu_erf_grad(bf16_hi(u.y)));
  o.z = pack_bf16x2(bf16_lo(a.z) * gelu_erf_grad(bf16_lo(u.z)), bf16_hi(a.z) * gelu_erf_grad(bf16_hi(u.z)));
  o.w = pack_bf16x2(bf16_lo(a.w) * gelu_erf_grad(bf16_lo(u.w)), bf16_hi(a.w) * gelu_erf_grad(bf16_hi(u.w)));
  *reinterpret_cast<uint4*>(dx + i) = o;
}

// Backward of NSP cross-entropy + seq_relationship Linear + pooler tanh, one thread per hidden column:
//   dlogit[b] = (softmax(logits[b]) - onehot(label[b])) * scale
//   dW[j,c] += sum_b dlogit[b,j] * pooled[b,c];  db[j] += sum_b dlogit[b,j]
//   dpre[b,c] = (dlogit[b,0] W[0,c] + dlogit[b,1] W[1,c]) * (1 - pooled[b,c]^2)
__global__ void __launch_bounds__(256) nsp_pool_bwd_kernel(const float* __restrict__ pooled,
                                                           const float* __restrict__ logits,
                                                           const int64_t* __restrict__ labels, int B,
                                                           const float* __restrict__ scale_dev,
                                                           const float* __restrict__ w, float* __restrict__ dw,
                                                           float* __restrict__ db, __nv_bfloat16* __restrict__ dpre) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kHidden) return;
  const float scale = __ldg(scale_dev);
  const float w0 = __ldg(w + c), w1 = __ldg(w + kHidden + c);
  float a0 = 0.f, a1 = 0.f, s0 = 0.f, s1 = 0.f;
  for (int b = 0; b < B; ++b) {
    const float l0 = __ldg(logits + 2 * b), l1 = __ldg(logits + 2 * b + 1);
    const float m = fmaxf(l0, l1);
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float inv = 1.0f / (e0 + e1);
    const int64_t lab = __ldg(labels + b);
    const float d0 = (e0 * inv - (lab == 0 ? 1.f : 0.f)) * scale;
    const float d1 = (e1 * inv - (lab == 1 ? 1.f : 0.f)) * scale;
    const float pv = __ldg(pooled + static_cast<int64_t>(b) * kHidden + c);
    a0 = fmaf(d0, pv, a0);
    a1 = fmaf(d1, pv, a1);
    s0 += d0;
    s1 += d1;
    dpre[static_cast<int64_t>(b) * kHidden + c] = __float2bfloat16_rn((d0 * w0 + d1 * w1) * (1.0f - pv * pv));
  }
  dw[c] += a0;
  dw[kHidden + c] += a1;
  if (c == 0) { db[0] += s0; db[1] += s1; }
}


// y[m, h*64 + d] = x[m, h*64 + d] * scales[h]: one thread per 8 columns (never crosses a head)
__global__ void __launch_bounds__(256) scale_heads_kernel(const __nv_bfloat16* __restrict__ x, int64_t n8,
                                                          const float* __restrict__ scales, __nv_bfloat16* __restrict__ y) {
  const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float s = __ldg(scales + (i % (kHidden / 8)) / 8);
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(x) + i);
  uint4 o;
  o.x = pack_bf16x2(bf16_lo(v.x) * s, bf16_hi(v.x) * s);
  o.y = pack_bf16x2(bf16_lo(v.y) * s, bf16_hi(v.y) * s);
  o.z = pack_bf16x2(bf16_lo(v.z) * s, bf16_hi(v.z) * s);
  o.w = pack_bf16x2(bf16_lo(v.w) * s, bf16_hi(v.w) * s);
  reinterpret_cast<uint4*>(y)[i] = o;
}

// Masked mean over the sequence: out[b, c] = sum_t mask[b, t] * x[b * seq_pad + t, c] / sum_t mask[b, t]  (t < seq_len).
// grid = (B, 3): block (b, k) owns 256 columns, its 8 warps stride over the tokens, a lane holds 8 columns (16 B loads).
__global__ void __launch_bounds__(256) masked_mean_pool_kernel(const __nv_bfloat16* __restrict__ x,
                                                               const int64_t* __restrict__ mask, int seq_len, int seq_pad,
                                                               float* __restrict__ out) {
  __shared__ float red[8][256];
  __shared__ int cnt[8];
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c0 = blockIdx.y * 256 + lane * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int n = 0;
  for (int t = warp; t < seq_len; t += 8) {
    if (mask != nullptr && __ldg(mask + static_cast<int64_t>(b) * seq_len + t) == 0) continue;   // warp-uniform
    ++n;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + (static_cast<int64_t>(b) * seq_pad + t) * kHidden + c0));
    acc[0] += bf16_lo(v.x); acc[1] += bf16_hi(v.x); acc[2] += bf16_lo(v.y); acc[3] += bf16_hi(v.y);
    acc[4] += bf16_lo(v.z); acc[5] += bf16_hi(v.z); acc[6] += bf16_lo(v.w); acc[7] += bf16_hi(v.w);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = acc[i];
  if (lane == 0) cnt[warp] = n;
  __syncthreads();
  float s = 0.f;
  int total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) { s += red[w][threadIdx.x]; total += cnt[w]; }
  out[static_cast<int64_t>(b) * kHidden + blockIdx.y * 256 + threadIdx.x] = s / static_cast<float>(total);   // 0 / 0 = NaN like torch
}

// dst[i] = float(src[i]) * scale : gradient bucket coming back from the bf16 all-reduce (mean = sum / world)
__global__ void unpack_scale_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst, int64_t n,
                                    float scale) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x * 8;
  for (int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 8; i < n; i += stride) {
    if (i + 8 <= n) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + i));
      float4* d = reinterpret_cast<float4*>(dst + i);
      d[0] = make_float4(bf16_lo(v.x) * scale, bf16_hi(v.x) * scale, bf16_lo(v.y) * scale, bf16_hi(v.y) * scale);
      d[1] = make_float4(bf16_lo(v.z) * scale, bf16_hi(v.z) * scale, bf16_lo(v.w) * scale, bf16_hi(v.w) * scale);
    } else {
      for (int64_t j = i; j < n; ++j) dst[j] = __bfloat162float(src[j]) * scale;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Sequence-classification head (fine-tuned checkpoints; reference stonkgs_finetuning.py:237-346):
//   logits = pooled . W^T + b   (W [L,768], L = num_labels <= 32),  single-label cross-entropy.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxLabels = 32;

// one warp per sample
__global__ void __launch_bounds__(256) cls_head_fwd_kernel(const float* __restrict__ pooled, int B, int L,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           const int64_t* __restrict__ labels,
                                                           float* __restrict__ logits, float* __restrict__ row_loss,
                                                           int* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  float x[kHidden / 32];
#pragma unroll
  for (int i = 0; i < kHidden / 32; ++i) x[i] = __ldg(pooled + static_cast<int64_t>(b) * kHidden + lane + 32 * i);
  float mine = 0.f;   // lane j keeps logit j
  for (int j = 0; j < L; ++j) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kHidden / 32; ++i) s = fmaf(x[i], __ldg(w + static_cast<int64_t>(j) * kHidden + lane + 32 * i), s);
    s = warp_sum(s) + __ldg(bias + j);
    if (lane == j) mine = s;
  }
  if (lane < L) logits[static_cast<int64_t>(b) * L + lane] = mine;
  if (labels && row_loss) {
    const float m = warp_max(lane < L ? mine : -INFINITY);
    const float e = lane < L ? expf(mine - m) : 0.f;
    const float lse = m + logf(warp_sum(e));
    const int64_t lab = __ldg(labels + b);
    if (lab < 0 || lab >= L) {
      if (lane == 0) { row_loss[b] = 0.f; if (err_flag) atomicOr(err_flag, 2); }
    } else {
      const float tgt = __shfl_sync(0xffffffffu, mine, static_cast<int>(lab));
      if (lane == 0) row_loss[b] = lse - tgt;
    }
  }
}

// dlogit[b, j] = (softmax(logits[b])[j] - [j == label[b]]) * scale   (one warp per sample)
__global__ void __launch_bounds__(256) cls_dlogit_kernel(const float* __restrict__ logits,
                                                         const int64_t* __restrict__ labels, int B, int L,
                                                         const float* __restrict__ scale_dev,
                                                         float* __restrict__ dlogit) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (b >= B) return;
  const float v = lane < L ? __ldg(logits + static_cast<int64_t>(b) * L + lane) : -INFINITY;
  const float m = warp_max(v);
  const float e = lane < L ? expf(v - m) : 0.f;
  const float inv = 1.0f / warp_sum(e);
  const int64_t lab = __ldg(labels + b);
  if (lane < L) dlogit[static_cast<int64_t>(b) * L + lane] = (e * inv - (lane == lab ? 1.f : 0.f)) * __ldg(scale_dev);
}

// one thread per hidden column c:
//   dW[j,c] += sum_b dlogit[b,j] pooled[b,c];  db[j] += sum_b dlogit[b,j]
//   dpre[b,c] = (sum_j dlogit[b,j] W[j,c]) * (1 - pooled[b,c]^2)      (backward of the pooler's tanh, HF:456-468)
__global__ void __launch_bounds__(256) cls_pool_bwd_kernel(const float* __restrict__ pooled,
                                                           const float* __restrict__ dlogit, int B, int L,
                                                           const float* __restrict__ w, float* __restrict__ dw,
                                                           float* __restrict__ db, __nv_bfloat16* __restrict__ dpre) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= kHidden) return;
  float wc[kMaxLabels], acc[kMaxLabels], sb[kMaxLabels];
#pragma unroll
  for (int j = 0; j < kMaxLabels; ++j) {
    wc[j] = j < L ? __ldg(w + static_cast<int64_t>(j) * kHidden + c) : 0.f;
    acc[j] = 0.f;
    sb[j] = 0.f;
  }
  for (int b = 0; b < B; ++b) {
    const float pv = __ldg(pooled + static_cast<int64_t>(b) * kHidden + c);
    float dp = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxLabels; ++j) {
      if (j < L) {
        const float d = __ldg(dlogit + static_cast<int64_t>(b) * L + j);   // same address across the warp: broadcast
        acc[j] = fmaf(d, pv, acc[j]);
        sb[j] += d;
        dp = fmaf(d, wc[j], dp);
      }
    }
    dpre[static_cast<int64_t>(b) * kHidden + c] = __float2bfloat16_rn(dp * (1.0f - pv * pv));
  }
#pragma unroll
  for (int j = 0; j < kMaxLabels; ++j) {
    if (j < L) {
      dw[static_cast<int64_t>(j) * kHidden + c] += acc[j];
      if (c == 0) db[j] += sb[j];
    }
  }
}

}  // namespace stk

using namespace stk;

#define STK_LAUNCHED()                                  \
  do {                                                  \
    STK_CHECK_CUDA(cudaGetLastError());                 \
    g_launches.fetch_add(1, std::memory_order_relaxed); \
    return STK_OK;                                      \
  } while (0)

extern "C" int stk_mask_to_bias(int device, void* stream, const int64_t* mask, int64_t n, float* bias) {
  STK_REQUIRE(mask && bias && n > 0, "stk_mask_to_bias: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  mask_to_bias_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(mask, n, bias);
  STK_LAUNCHED();
}

extern "C" int stk_cast_f32_to_bf16(int device, void* stream, const float* src, void* dst, int64_t n) {
  STK_REQUIRE(src && dst && n > 0, "stk_cast_f32_to_bf16: bad arguments");
  STK_REQUIRE((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "stk_cast_f32_to_bf16: pointers must be 16-byte aligned");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms(device)) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_f32_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src, static_cast<__nv_bfloat16*>(dst), n);
  STK_LAUNCHED();
}

extern "C" int stk_gather_rows(int device, void* stream, const void* src, const int32_t* idx, int n_rows, void* dst) {
  STK_REQUIRE(src && idx && dst && n_rows > 0, "stk_gather_rows: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  gather_rows_kernel<<<(n_rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), idx, n_rows, static_cast<__nv_bfloat16*>(dst));
  STK_LAUNCHED();
}

extern "C" int stk_scatter_add_rows(int device, void* stream, const void* src, const int32_t* idx, int n_rows,
                                    void* dst) {
  STK_REQUIRE(src && idx && dst && n_rows > 0, "stk_scatter_add_rows: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  scatter_add_rows_kernel<<<(n_rows + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src), idx, n_rows, static_cast<__nv_bfloat16*>(dst));
  STK_LAUNCHED();
}

extern "C" int stk_colsum(int device, void* stream, const void* x, int64_t ld, int M, int N, float* out,
                          int accumulate) {
  STK_REQUIRE(x && out && M > 0 && N > 0 && ld % 2 == 0 && N % 2 == 0, "stk_colsum: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (!accumulate) STK_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, s));
  const int strips = (N + 63) / 64;
  int slabs = (num_sms(device) * 4 + strips - 1) / strips;
  int rows_per_block = (M + slabs - 1) / slabs;
  if (rows_per_block < 64) rows_per_block = 64;
  slabs = (M + rows_per_block - 1) / rows_per_block;
  colsum_kernel<<<dim3(strips, slabs), 256, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ld, M, N, rows_per_block, out);
  STK_LAUNCHED();
}

extern "C" int stk_ce_finalize(int device, void* stream, const float* ce_partial, int64_t ce_pitch,
                               const float* tgt_logit, int M, float* lse, float* row_loss) {
  STK_REQUIRE(ce_partial && lse && M > 0 && ce_pitch > 0, "stk_ce_finalize: bad arguments");
  STK_REQUIRE(row_loss == nullptr || tgt_logit != nullptr, "stk_ce_finalize: row_loss needs tgt_logit");
  STK_CHECK_CUDA(cudaSetDevice(device));
  ce_finalize_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float2*>(ce_partial), ce_pitch, tgt_logit, M, lse, row_loss);
  STK_LAUNCHED();
}

extern "C" int stk_nsp_head_fwd(int device, void* stream, const float* pooled, int B, const float* w, const float* b,
                                const int64_t* labels, float* logits, float* row_loss, int* err_flag) {
  STK_REQUIRE(pooled && w && b && logits && B > 0, "stk_nsp_head_fwd: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  nsp_head_fwd_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pooled, B, w, b, labels, logits,
                                                                                   row_loss, err_flag);
  STK_LAUNCHED();
}

extern "C" int stk_gelu_bwd(int device, void* stream, const void* dy, const void* pre, int64_t n, void* dx) {
  STK_REQUIRE(dy && pre && dx && n > 0 && n % 8 == 0, "stk_gelu_bwd: bad arguments (n must be a multiple of 8)");
  STK_CHECK_CUDA(cudaSetDevice(device));
  gelu_bwd_kernel<<<static_cast<unsigned>((n / 8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(pre), n, static_cast<__nv_bfloat16*>(dx));
  STK_LAUNCHED();
}

extern "C" int stk_nsp_pool_bwd(int device, void* stream, const float* pooled, const float* logits,
                                const int64_t* labels, int B, const float* scale_dev, const float* w, float* dw,
                                float* db, void* dpre) {
  STK_REQUIRE(pooled && logits && labels && scale_dev && w && dw && db && dpre && B > 0, "stk_nsp_pool_bwd: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  nsp_pool_bwd_kernel<<<3, 256, 0, static_cast<cudaStream_t>(stream)>>>(pooled, logits, labels, B, scale_dev, w, dw, db,
                                                                      static_cast<__nv_bfloat16*>(dpre));
  STK_LAUNCHED();
}

extern "C" int stk_masked_mean_pool(int device, void* stream, const void* x, const int64_t* mask, int B, int seq_len,
                                    int seq_pad, float* out) {
  STK_REQUIRE(x && out && B > 0 && seq_len > 0 && seq_pad >= seq_len, "stk_masked_mean_pool: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  masked_mean_pool_kernel<<<dim3(B, kHidden / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), mask, seq_len, seq_pad, out);
  STK_LAUNCHED();
}

extern "C" int stk_scale_heads(int device, void* stream, const void* x, int M, const float* scales, void* y) {
  STK_REQUIRE(x && y && scales && M > 0, "stk_scale_heads: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  const int64_t n8 = static_cast<int64_t>(M) * (kHidden / 8);
  scale_heads_kernel<<<static_cast<unsigned>((n8 + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), n8, scales, static_cast<__nv_bfloat16*>(y));
  STK_LAUNCHED();
}

extern "C" int stk_unpack_scale(int device, void* stream, const void* src_bf16, float* dst, int64_t n, float scale) {
  STK_REQUIRE(src_bf16 && dst && n > 0, "stk_unpack_scale: bad arguments");
  STK_REQUIRE((reinterpret_cast<uintptr_t>(src_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0,
              "stk_unpack_scale: pointers must be 16-byte aligned");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int64_t blocks = (n / 8 + 255) / 256;
  const int64_t cap = static_cast<int64_t>(num_sms(device)) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  unpack_scale_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(src_bf16), dst, n, scale);
  STK_LAUNCHED();
}

extern "C" int stk_cls_head_fwd(int device, void* stream, const float* pooled, int B, int num_labels, const float* w,
                                const float* b, const int64_t* labels, float* logits, float* row_loss, int* err_flag) {
  STK_REQUIRE(pooled && w && b && logits && B > 0, "stk_cls_head_fwd: bad arguments");
  STK_REQUIRE(num_labels >= 1 && num_labels <= kMaxLabels, "stk_cls_head_fwd: num_labels must be in [1, 32] (got %d)", num_labels);
  STK_CHECK_CUDA(cudaSetDevice(device));
  cls_head_fwd_kernel<<<(B + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(pooled, B, num_labels, w, b, labels,
                                                                                 logits, row_loss, err_flag);
  STK_LAUNCHED();
}

extern "C" int stk_cls_pool_bwd(int device, void* stream, const float* pooled, const float* logits,
                                const int64_t* labels, int B, int num_labels, const float* scale_dev, const float* w,
                                float* dlogit_ws, float* dw, float* db, void* dpre_bf16) {
  STK_REQUIRE(pooled && logits && labels && scale_dev && w && dlogit_ws && dw && db && dpre_bf16 && B > 0,
              "stk_cls_pool_bwd: bad arguments");
  STK_REQUIRE(num_labels >= 1 && num_labels <= kMaxLabels, "stk_cls_pool_bwd: num_labels must be in [1, 32] (got %d)", num_labels);
  STK_CHECK_CUDA(cudaSetDevice(device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cls_dlogit_kernel<<<(B + 7) / 8, 256, 0, st>>>(logits, labels, B, num_labels, scale_dev, dlogit_ws);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cls_pool_bwd_kernel<<<3, 256, 0, st>>>(pooled, dlogit_ws, B, num_labels, w, dw, db,
                                          static_cast<__nv_bfloat16*>(dpre_bf16));
  STK_LAUNCHED();
}
