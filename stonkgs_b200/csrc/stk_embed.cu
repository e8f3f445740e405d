// stk_embed.cu — HBM-bound row kernels: the two fused embedding stages and LayerNorm (fwd + bwd).
//
// All of them are one-warp-per-row streaming kernels over rows of 768 elements: every lane owns
// 24 elements as six 4-element chunks (chunk = lane + 32*i), so each warp-wide access is a fully
// coalesced 512 B (fp32 rows) or 256 B (bf16 rows) segment; statistics stay in fp32 registers and
// the row is read exactly once.  Algorithmic traffic per token: 768*4 B (fp32 source row) or
// 768*2 B (bf16 source row) read + 768*2 B bf16 written (4 608 B/token).
//
// Reference lines: HF modeling_bert.py:72-112 (BertEmbeddings.forward: (x + type) + pos, LayerNorm
// eps 1e-12); stonkgs_model.py:178 (ids-only LM backbone call), :182-200 (KG dict lookup, concat,
// cast), :204-210 (joint encoder on inputs_embeds).
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"
#include "stk_rng.cuh"

namespace stk {

extern std::atomic<long long> g_launches;

constexpr int kChunks = 6;       // 4-element chunks per lane
constexpr int kPerLane = 24;     // elements per lane
constexpr int kRowWarps = 8;     // rows (warps) per 256-thread block

__device__ __forceinline__ void load_f32_row(const float* row, int lane, float (&v)[kPerLane]) {
  const float4* p = reinterpret_cast<const float4*>(row);
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const float4 t = __ldg(p + lane + 32 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void load_bf16_row(const __nv_bfloat16* row, int lane, float (&v)[kPerLane]) {
  const uint2* p = reinterpret_cast<const uint2*>(row);
#pragma unroll
  for (int i = 0; i < kChunks; ++i) {
    const uint2 t = __ldg(p + lane + 32 * i);
    v[4 * i] = bf16_lo(t.x); v[4 * i + 1] = bf16_hi(t.x); v[4 * i + 2] = bf16_lo(t.y); v[4 * i + 3] = bf16_hi(t.y);
  }
}
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* row, int lane, const float (&v)[kPerLane]) {
  uint2* p = reinterpret_cast<uint2*>(row);
#pragma unroll
  for (int i = 0; i < kChunks; ++i)
    p[lane + 32 * i] = make_uint2(pack_bf16x2(v[4 * i], v[4 * i + 1]), pack_bf16x2(v[4 * i + 2], v[4 * i + 3]));
}
__device__ __forceinline__ void store_f32_row(float* row, int lane, const float (&v)[kPerLane]) {
  float4* p = reinterpret_cast<float4*>(row);
#pragma unroll
  for (int i = 0; i < kChunks; ++i) p[lane + 32 * i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}

// mean / rstd of one 768-wide row held across a warp (two-pass, fp32)
__device__ __forceinline__ void row_stats(const float (&v)[kPerLane], float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) s += v[i];
  mean = warp_sum(s) * (1.0f / kHidden);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    const float d = v[i] - mean;
    q = fmaf(d, d, q);
  }
  rstd = rsqrtf(warp_sum(q) * (1.0f / kHidden) + kLnEps);
}
__device__ __forceinline__ void normalize(float (&v)[kPerLane], float mean, float rstd, const float* gamma,
                                          const float* beta, int lane) {
  float g[kPerLane], b[kPerLane];
  load_f32_row(gamma, lane, g);
  load_f32_row(beta, lane, b);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
}

// ------------------------------------------------------------------------------------------------
// A1: out = LN(word[id] + type[0] + pos[p])
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_text_ln_kernel(const int64_t* __restrict__ ids, int64_t ids_pitch, int B,
                                                            int S, const float* __restrict__ word, int vocab,
                                                            const float* __restrict__ pos,
                                                            const float* __restrict__ type_emb,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ out, int* err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (tok >= static_cast<int64_t>(B) * S) return;
  const int b = static_cast<int>(tok / S), t = static_cast<int>(tok % S);
  int64_t id = __ldg(ids + b * ids_pitch + t);
  if (id < 0 || id >= vocab) {
    if (lane == 0 && err_flag) atomicOr(err_flag, 1);
    id = 0;
  }
  float v[kPerLane], a[kPerLane];
  load_f32_row(word + id * kHidden, lane, v);
  load_f32_row(type_emb, lane, a);  // token type 0
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] += a[i];
  load_f32_row(pos + static_cast<int64_t>(t) * kHidden, lane, a);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] += a[i];
  float mean, rstd;
  row_stats(v, mean, rstd);
  normalize(v, mean, rstd, gamma, beta, lane);
  store_bf16_row(out + tok * kHidden, lane, v);
}

// ------------------------------------------------------------------------------------------------
// A3+A4: joint stage. Gathered source row (LM hidden state or KG table row) + type + pos -> LN.
// ------------------------------------------------------------------------------------------------
// Shape of the joint sequence: T text tokens (LM-backbone rows) followed by S - T KG tokens; activations use SP >= S
// rows per pair (SP = S for STonKGs: 256 + 256; the TransE variant, transestonkgs_model.py:44,93, has S = 260 = 256 + 4
// and runs the encoder on SP = 384 rows so that the attention kernels see whole 128-key blocks; rows t >= S are zero
// and masked out as keys).
struct JointShape {
  int T, S, SP;
};

__device__ __forceinline__ bool joint_source_row(const int64_t* input_ids, const int64_t* token_type_ids, int64_t b, int t,
                                                 JointShape sh, const __nv_bfloat16* lm_hidden, const float* kg_table,
                                                 int64_t table_rows, int lane, float (&v)[kPerLane], int& tt) {
  bool ok = true;
  if (t < sh.T) {
    load_bf16_row(lm_hidden + (b * sh.T + t) * kHidden, lane, v);
  } else {
    int64_t id = __ldg(input_ids + b * sh.S + t);
    if (id < 0 || id >= table_rows) { ok = false; id = 0; }
    load_f32_row(kg_table + id * kHidden, lane, v);
  }
  tt = token_type_ids ? static_cast<int>(__ldg(token_type_ids + b * sh.S + t)) : (t >= sh.T ? 1 : 0);
  if (tt < 0 || tt > 1) { ok = false; tt = 0; }
  return ok;
}

__global__ void __launch_bounds__(256)
embed_joint_ln_kernel(const int64_t* __restrict__ input_ids, const int64_t* __restrict__ token_type_ids, int B,
                      JointShape sh, const __nv_bfloat16* __restrict__ lm_hidden, const float* __restrict__ kg_table,
                      int64_t table_rows, const float* __restrict__ pos, const float* __restrict__ type_emb,
                      const float* __restrict__ gamma, const float* __restrict__ beta,
                      __nv_bfloat16* __restrict__ out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                      float* __restrict__ inputs_embeds_out, int* err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t tok = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);   // row of the padded layout
  if (tok >= static_cast<int64_t>(B) * sh.SP) return;
  const int64_t b = tok / sh.SP;
  const int t = static_cast<int>(tok - b * sh.SP);
  float v[kPerLane], a[kPerLane];
  if (t >= sh.S) {   // padding row of the TransE layout: finite (zero) activations, never attended to
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) v[i] = 0.f;
    if (inputs_embeds_out) store_f32_row(inputs_embeds_out + tok * kHidden, lane, v);
    if (mean_out && lane == 0) { mean_out[tok] = 0.f; rstd_out[tok] = 0.f; }
    store_bf16_row(out + tok * kHidden, lane, v);
    return;
  }
  int tt;
  if (!joint_source_row(input_ids, token_type_ids, b, t, sh, lm_hidden, kg_table, table_rows, lane, v, tt)) {
    if (lane == 0 && err_flag) atomicOr(err_flag, 1);
  }
  if (inputs_embeds_out) store_f32_row(inputs_embeds_out + tok * kHidden, lane, v);
  load_f32_row(type_emb + tt * kHidden, lane, a);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] += a[i];
  load_f32_row(pos + static_cast<int64_t>(t) * kHidden, lane, a);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) v[i] += a[i];
  float mean, rstd;
  row_stats(v, mean, rstd);
  if (mean_out && lane == 0) { mean_out[tok] = mean; rstd_out[tok] = rstd; }
  normalize(v, mean, rstd, gamma, beta, lane);
  store_bf16_row(out + tok * kHidden, lane, v);
}

// LayerNorm backward for one row held across a warp.  On return dy[] holds dx, xh[] holds x_hat.
__device__ __forceinline__ void ln_row_bwd(float (&dy)[kPerLane], float (&xh)[kPerLane], const float (&g)[kPerLane],
                                           float mean, float rstd, float (&dg_acc)[kPerLane],
                                           float (&db_acc)[kPerLane]) {
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) {
    xh[i] = (xh[i] - mean) * rstd;
    dg_acc[i] = fmaf(dy[i], xh[i], dg_acc[i]);
    db_acc[i] += dy[i];
    dy[i] *= g[i];
    s1 += dy[i];
    s2 = fmaf(dy[i], xh[i], s2);
  }
  s1 = warp_sum(s1) * (1.0f / kHidden);
  s2 = warp_sum(s2) * (1.0f / kHidden);
#pragma unroll
  for (int i = 0; i < kPerLane; ++i) dy[i] = (dy[i] - s1 - xh[i] * s2) * rstd;
}

// Backward of the joint stage: one block per sequence position t (so dpos[t] is owned by the block),
// warps stride over the batch; per-block partial sums for dtype/dgamma/dbeta go out as atomics.
__global__ void __launch_bounds__(128)
embed_joint_ln_bwd_kernel(const int64_t* __restrict__ input_ids, const int64_t* __restrict__ token_type_ids, int B,
                          JointShape sh, const __nv_bfloat16* __restrict__ lm_hidden, const float* __restrict__ kg_table,
                          int64_t table_rows, const float* __restrict__ pos, const float* __restrict__ type_emb,
                          const float* __restrict__ gamma, const float* __restrict__ mean_in,
                          const float* __restrict__ rstd_in, const __nv_bfloat16* __restrict__ dy_in,
                          float* __restrict__ dpos, float* __restrict__ dtype, float* __restrict__ dgamma,
                          float* __restrict__ dbeta) {
  __shared__ float red[4][kHidden];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int t = blockIdx.x;
  float g[kPerLane], pz[kPerLane];
  load_f32_row(gamma, lane, g);
  load_f32_row(pos + static_cast<int64_t>(t) * kHidden, lane, pz);
  float dg[kPerLane] = {}, db[kPerLane] = {}, dp[kPerLane] = {}, dt0[kPerLane] = {}, dt1[kPerLane] = {};
  for (int b = warp; b < B; b += 4) {
    const int64_t tok = static_cast<int64_t>(b) * sh.SP + t;
    float x[kPerLane], a[kPerLane], dy[kPerLane];
    int tt;
    joint_source_row(input_ids, token_type_ids, b, t, sh, lm_hidden, kg_table, table_rows, lane, x, tt);
    load_f32_row(type_emb + tt * kHidden, lane, a);
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) x[i] = (x[i] + a[i]) + pz[i];
    load_bf16_row(dy_in + tok * kHidden, lane, dy);
    ln_row_bwd(dy, x, g, __ldg(mean_in + tok), __ldg(rstd_in + tok), dg, db);
#pragma unroll
    for (int i = 0; i < kPerLane; ++i) {
      dp[i] += dy[i];
      if (tt == 0) dt0[i] += dy[i]; else dt1[i] += dy[i];
    }
  }
  // block reduction of the five 768-vectors, one at a time through shared memory
  auto reduce_out = [&](float (&acc)[kPerLane], float* dst, bool atomic) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kChunks; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][(lane + 32 * i) * 4 + j] = acc[4 * i + j];
    __syncthreads();
    for (int c = threadIdx.x; c < kHidden; c += 128) {
      const float s = red[0][c] + red[1][c] + red[2][c] + red[3][c];
      if (atomic) atomicAdd(dst + c, s); else dst[c] += s;
    }
  };
  reduce_out(dp, dpos + static_cast<int64_t>(t) * kHidden, false);
  reduce_out(dt0, dtype, true);
  reduce_out(dt1, dtype + kHidden, true);
  reduce_out(dg, dgamma, true);
  reduce_out(db, dbeta, true);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over bf16 rows
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, int M,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta,
                                                            __nv_bfloat16* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + (threadIdx.x >> 5);
  if (row >= M) return;
  float v[kPerLane];
  load_bf16_row(x + row * kHidden, lane, v);
  float mean, rstd;
  row_stats(v, mean, rstd);
  if (mean_out && lane == 0) { mean_out[row] = mean; rstd_out[row] = rstd; }
  normalize(v, mean, rstd, gamma, beta, lane);
  store_bf16_row(y + row * kHidden, lane, v);
}

// grid-stride over rows; per-warp register accumulators for dgamma/dbeta, block reduce, atomics out
template <bool FUSED>
__global__ void __launch_bounds__(256)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy_in, const __nv_bfloat16* __restrict__ x_in, int M,
                     const float* __restrict__ gamma, const float* __restrict__ mean_in,
                     const float* __restrict__ rstd_in, __nv_bfloat16* __restrict__ dx_out, float* __restrict__ dgamma,
                     float* __restrict__ dbeta, __nv_bfloat16* __restrict__ dxm_out, float* __restrict__ dbias,
                     uint32_t seed, uint32_t site, uint32_t thr) {
  __shared__ float red[kRowWarps][kHidden];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[kPerLane];
  load_f32_row(gamma, lane, g);
  float dg[kPerLane] = {}, db[kPerLane] = {};
  float dbs[FUSED ? kPerLane : 1] = {};
  const uint32_t thr4 = drop_thr4(thr);
  const float dscale = drop_scale(thr);
  for (int64_t row = static_cast<int64_t>(blockIdx.x) * kRowWarps + warp; row < M;
       row += static_cast<int64_t>(gridDim.x) * kRowWarps) {
    float dy[kPerLane], x[kPerLane];
    load_bf16_row(dy_in + row * kHidden, lane, dy);
    load_bf16_row(x_in + row * kHidden, lane, x);
    ln_row_bwd(dy, x, g, __ldg(mean_in + row), __ldg(rstd_in + row), dg, db);
    store_bf16_row(dx_out + row * kHidden, lane, dy);
    if (FUSED) {
      // what the dense layer before this LayerNorm sees: the gradient through its dropout mask (the residual branch
      // takes dx as it is), plus that layer's bias gradient = column sum of the same rows
      if (dxm_out) {
        const uint32_t row_key = drop_row_key(seed, site, static_cast<uint32_t>(row));
#pragma unroll
        for (int i = 0; i < kChunks; ++i) {
          const uint32_t chunk = static_cast<uint32_t>(lane + 32 * i);   // columns 4*chunk .. 4*chunk + 3
          uint32_t w0, w1;
          drop_words(row_key, chunk >> 1, w0, w1);
          const uint32_t signs = drop_signs((chunk & 1u) ? w1 : w0, thr4);
#pragma unroll
          for (int k = 0; k < 4; ++k) dy[4 * i + k] = drop_keep(signs, k) ? dy[4 * i + k] * dscale : 0.f;
        }
        store_bf16_row(dxm_out + row * kHidden, lane, dy);
      }
      if (dbias) {
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) dbs[i] += dy[i];
      }
    }
  }
  auto reduce_out = [&](float (&acc)[kPerLane], float* dst) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kChunks; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[warp][(lane + 32 * i) * 4 + j] = acc[4 * i + j];
    __syncthreads();
    for (int c = threadIdx.x; c < kHidden; c += 256) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) s += red[w][c];
      atomicAdd(dst + c, s);
    }
  };
  reduce_out(dg, dgamma);
  reduce_out(db, dbeta);
  if constexpr (FUSED) {
    if (dbias) reduce_out(dbs, dbias);
  }
}

}  // namespace stk

using namespace stk;

#define STK_LAUNCHED()                                         \
  do {                                                         \
    STK_CHECK_CUDA(cudaGetLastError());                        \
    g_launches.fetch_add(1, std::memory_order_relaxed);        \
    return STK_OK;                                             \
  } while (0)

extern "C" int stk_embed_text_ln_fwd(int device, void* stream, const int64_t* ids, int64_t ids_pitch, int B, int S,
                                     const float* word, int vocab, const float* pos, const float* type_emb,
                                     const float* gamma, const float* beta, void* out, int* err_flag) {
  STK_REQUIRE(B > 0 && S > 0 && S <= 512, "stk_embed_text_ln_fwd: bad shape B=%d S=%d", B, S);
  STK_REQUIRE(ids && word && pos && type_emb && gamma && beta && out, "stk_embed_text_ln_fwd: null pointer");
  STK_CHECK_CUDA(cudaSetDevice(device));
  const int64_t rows = static_cast<int64_t>(B) * S;
  embed_text_ln_kernel<<<static_cast<unsigned>((rows + kRowWarps - 1) / kRowWarps), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(ids, ids_pitch, B, S, word, vocab, pos, type_emb, gamma,
                                                              beta, static_cast<__nv_bfloat16*>(out), err_flag);
  STK_LAUNCHED();
}

static int check_joint_shape(const char* who, int B, int T, int S, int SP) {
  STK_REQUIRE(B > 0, "%s: bad batch %d", who, B);
  STK_REQUIRE(T > 0 && S > T && SP >= S, "%s: bad joint shape (text %d, sequence %d, padded %d)", who, T, S, SP);
  return STK_OK;
}

extern "C" int stk_embed_joint_ln_fwd_shape(int device, void* stream, const int64_t* input_ids,
                                            const int64_t* token_type_ids, int B, int text_len, int seq_len, int seq_pad,
                                            const void* lm_hidden, const float* kg_table, int64_t table_rows,
                                            const float* pos, const float* type_emb, const float* gamma,
                                            const float* beta, void* out, float* mean, float* rstd,
                                            float* inputs_embeds_out, int* err_flag) {
  if (int rc = check_joint_shape("stk_embed_joint_ln_fwd", B, text_len, seq_len, seq_pad)) return rc;
  STK_REQUIRE(input_ids && lm_hidden && kg_table && pos && type_emb && gamma && beta && out,
              "stk_embed_joint_ln_fwd: null pointer");
  STK_REQUIRE((mean == nullptr) == (rstd == nullptr), "stk_embed_joint_ln_fwd: mean/rstd must both be given or both NULL");
  STK_CHECK_CUDA(cudaSetDevice(device));
  const int64_t rows = static_cast<int64_t>(B) * seq_pad;
  embed_joint_ln_kernel<<<static_cast<unsigned>((rows + kRowWarps - 1) / kRowWarps), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      input_ids, token_type_ids, B, JointShape{text_len, seq_len, seq_pad}, static_cast<const __nv_bfloat16*>(lm_hidden),
      kg_table, table_rows, pos, type_emb, gamma, beta, static_cast<__nv_bfloat16*>(out), mean, rstd, inputs_embeds_out,
      err_flag);
  STK_LAUNCHED();
}

extern "C" int stk_embed_joint_ln_fwd(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                                      int B, const void* lm_hidden, const float* kg_table, int64_t table_rows,
                                      const float* pos, const float* type_emb, const float* gamma, const float* beta,
                                      void* out, float* mean, float* rstd, float* inputs_embeds_out, int* err_flag) {
  return stk_embed_joint_ln_fwd_shape(device, stream, input_ids, token_type_ids, B, 256, 512, 512, lm_hidden, kg_table,
                                      table_rows, pos, type_emb, gamma, beta, out, mean, rstd, inputs_embeds_out, err_flag);
}

extern "C" int stk_embed_joint_ln_bwd_shape(int device, void* stream, const int64_t* input_ids,
                                            const int64_t* token_type_ids, int B, int text_len, int seq_len, int seq_pad,
                                            const void* lm_hidden, const float* kg_table, int64_t table_rows,
                                            const float* pos, const float* type_emb, const float* gamma,
                                            const float* mean, const float* rstd, const void* dy, float* dpos,
                                            float* dtype, float* dgamma, float* dbeta) {
  if (int rc = check_joint_shape("stk_embed_joint_ln_bwd", B, text_len, seq_len, seq_pad)) return rc;
  STK_REQUIRE(input_ids && lm_hidden && kg_table && pos && type_emb && gamma && mean && rstd && dy && dpos && dtype &&
                  dgamma && dbeta,
              "stk_embed_joint_ln_bwd: null pointer");
  STK_CHECK_CUDA(cudaSetDevice(device));
  embed_joint_ln_bwd_kernel<<<seq_len, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      input_ids, token_type_ids, B, JointShape{text_len, seq_len, seq_pad}, static_cast<const __nv_bfloat16*>(lm_hidden),
      kg_table, table_rows, pos, type_emb, gamma, mean, rstd, static_cast<const __nv_bfloat16*>(dy), dpos, dtype, dgamma,
      dbeta);
  STK_LAUNCHED();
}

extern "C" int stk_embed_joint_ln_bwd(int device, void* stream, const int64_t* input_ids, const int64_t* token_type_ids,
                                      int B, const void* lm_hidden, const float* kg_table, int64_t table_rows,
                                      const float* pos, const float* type_emb, const float* gamma, const float* mean,
                                      const float* rstd, const void* dy, float* dpos, float* dtype, float* dgamma,
                                      float* dbeta) {
  return stk_embed_joint_ln_bwd_shape(device, stream, input_ids, token_type_ids, B, 256, 512, 512, lm_hidden, kg_table,
                                      table_rows, pos, type_emb, gamma, mean, rstd, dy, dpos, dtype, dgamma, dbeta);
}

extern "C" int stk_layernorm_fwd(int device, void* stream, const void* x, int M, const float* gamma, const float* beta,
                                 void* y, float* mean, float* rstd) {
  STK_REQUIRE(M > 0 && x && y && gamma && beta, "stk_layernorm_fwd: bad arguments");
  STK_REQUIRE((mean == nullptr) == (rstd == nullptr), "stk_layernorm_fwd: mean/rstd must both be given or both NULL");
  STK_CHECK_CUDA(cudaSetDevice(device));
  layernorm_fwd_kernel<<<(M + kRowWarps - 1) / kRowWarps, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x), M, gamma, beta, static_cast<__nv_bfloat16*>(y), mean, rstd);
  STK_LAUNCHED();
}

extern "C" int stk_layernorm_bwd(int device, void* stream, const void* dy, const void* x, int M, const float* gamma,
                                 const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta) {
  STK_REQUIRE(M > 0 && dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "stk_layernorm_bwd: bad arguments");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int grid = (M + kRowWarps - 1) / kRowWarps;
  const int cap = num_sms(device) * 4;
  if (grid > cap) grid = cap;
  layernorm_bwd_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), M, gamma, mean, rstd,
      static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, nullptr, nullptr, 0u, 0u, 0u);
  STK_LAUNCHED();
}

extern "C" int stk_layernorm_bwd_fused(int device, void* stream, const void* dy, const void* x, int M, const float* gamma,
                                       const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                                       void* dxm, float* dbias, uint32_t seed, uint32_t site, uint32_t thr) {
  STK_REQUIRE(M > 0 && dy && x && gamma && mean && rstd && dx && dgamma && dbeta, "stk_layernorm_bwd_fused: bad arguments");
  STK_REQUIRE(thr < 128 && (dxm != nullptr || thr == 0), "stk_layernorm_bwd_fused: dropout needs the dxm output and thr < 128");
  STK_CHECK_CUDA(cudaSetDevice(device));
  int grid = (M + kRowWarps - 1) / kRowWarps;
  const int cap = num_sms(device) * 4;
  if (grid > cap) grid = cap;
  layernorm_bwd_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(dy), static_cast<const __nv_bfloat16*>(x), M, gamma, mean, rstd,
      static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, thr > 0 ? static_cast<__nv_bfloat16*>(dxm) : nullptr, dbias, seed, site,
      thr);
  STK_LAUNCHED();
}
