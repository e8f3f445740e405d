// stk_heads.cu — the MLM / ELM heads of the pre-training step as single C-ABI calls:
//
//   stk_compact_labels   label selection of stonkgs_model.py:229-245 (labels != -100) on the device: row indices and
//                        int32 labels of the labelled positions at a FIXED capacity, so that the host never reads a
//                        count back (no torch.nonzero sync inside the training step)
//   stk_linear_ce_fwd    text_decoder / entity_decoder (stonkgs_model.py:62-73, no bias) + mean cross-entropy
//                        (:229-245) over the labelled rows: one tcgen05 GEMM whose epilogue keeps per-row
//                        (max, sum exp) per 128-column slab and the target logit — the [rows, V] logits never exist
//   stk_linear_ce_bwd    its backward, owning the vocabulary-chunk loop: per chunk (dlogit workspace <= 48 MB, L2
//                        resident) recompute the logits tile -> (softmax - onehot) * scale in the GEMM epilogue ->
//                        dT += dlogit W_chunk and dW_chunk += dlogit^T T
//   stk_query_workspace  workspace bytes of the calls that need one
//
// The GEMMs are stk_gemm (stk_gemm.cu); this file is host orchestration plus three small kernels.
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

extern std::atomic<long long> g_launches;

// One block walks the [B, width] label matrix in row-major order (the order torch.nonzero would give), so the
// compaction is deterministic.  Entries [count, capacity) are padding: row -1 (gathers as zeros, never scattered)
// and label -1 (no loss, zero gradient).
__global__ void __launch_bounds__(1024)
compact_labels_kernel(const int64_t* __restrict__ labels, int B, int width, int row_pitch, int col_offset, int vocab,
                      int capacity, int32_t* __restrict__ rows_out, int32_t* __restrict__ labels_out,
                      int32_t* __restrict__ count_out, int* __restrict__ err_flag) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_base = 0;
  __syncthreads();
  const int n = B * width;
  bool bad = false, overflow = false;
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    int64_t lab = -100;
    if (i < n) lab = __ldg(labels + i);
    bool valid = lab != -100;
    if (valid && (lab < 0 || lab >= vocab)) {   // torch's cross-entropy raises on such a target: flag it, skip the row
      bad = true;
      valid = false;
    }
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) s_warp[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll 8
    for (int w = 0; w < 32; ++w) {
      const int c = s_warp[w];
      before += w < warp ? c : 0;
      total += c;
    }
    const int pos = s_base + before + __popc(m & ((1u << lane) - 1u));
    if (valid) {
      if (pos < capacity) {
        const int b = i / width, t = i - b * width;
        rows_out[pos] = b * row_pitch + col_offset + t;
        labels_out[pos] = static_cast<int32_t>(lab);
      } else {
        overflow = true;
      }
    }
    __syncthreads();
    if (tid == 0) s_base += total;
    __syncthreads();
  }
  const int count = s_base;
  for (int i = min(count, capacity) + tid; i < capacity; i += 1024) {
    rows_out[i] = -1;
    labels_out[i] = -1;
  }
  if (tid == 0) count_out[0] = count;
  if (err_flag) {
    if (bad) atomicOr(err_flag, 2);
    if (overflow) atomicOr(err_flag, 4);
  }
}

// one warp per row: combine the (max, sumexp) slab partials into lse and the row loss (0 for padding rows)
__global__ void __launch_bounds__(256)
ce_rows_kernel(const float2* __restrict__ part, int64_t pitch, const float* __restrict__ tgt,
               const int32_t* __restrict__ labels, int M, float* __restrict__ lse_out, float* __restrict__ row_loss) {
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
    const bool on = labels == nullptr || __ldg(labels + row) >= 0;
    if (row_loss) row_loss[row] = on ? lse - __ldg(tgt + row) : 0.f;   // tgt stays NaN if no column matched the label
  }
}

// loss_count[0] = sum(row_loss) / count, loss_count[1] = count   (count = rows with label >= 0; 0 / 0 = NaN like torch)
__global__ void __launch_bounds__(1024)
ce_mean_kernel(const float* __restrict__ row_loss, const int32_t* __restrict__ labels, int M,
               float* __restrict__ loss_count) {
  __shared__ float s_sum[32];
  __shared__ int s_cnt[32];
  float s = 0.f;
  int c = 0;
  for (int i = threadIdx.x; i < M; i += 1024) {
    const bool on = labels == nullptr || __ldg(labels + i) >= 0;
    if (on) { s += __ldg(row_loss + i); ++c; }
  }
  s = warp_sum(s);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) { s_sum[threadIdx.x >> 5] = s; s_cnt[threadIdx.x >> 5] = c; }
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum(s_sum[threadIdx.x]);
    c = __reduce_add_sync(0xffffffffu, s_cnt[threadIdx.x]);
    if (threadIdx.x == 0) {
      loss_count[0] = s / static_cast<float>(c);
      loss_count[1] = static_cast<float>(c);
    }
  }
}

static inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

// vocabulary columns per backward chunk: the bf16 dlogit workspace [R, C] stays L2-resident (<= 48 MB)
static int ce_chunk_cols(int R, int V) {
  int64_t c = ((48ll << 20) / (2ll * (R > 0 ? R : 1))) / 256 * 256;
  if (c > 32768) c = 32768;
  if (c < 256) c = 256;
  const int64_t vmax = round_up(V, 256);
  return static_cast<int>(c < vmax ? c : vmax);
}

}  // namespace stk

using namespace stk;

#define STK_COUNT_LAUNCH()                              \
  do {                                                  \
    STK_CHECK_CUDA(cudaGetLastError());                 \
    g_launches.fetch_add(1, std::memory_order_relaxed); \
  } while (0)

extern "C" int64_t stk_query_workspace(int op, int64_t a, int64_t b) {
  switch (op) {
    case STK_WS_ATTN_BWD:        // a = B, b = S: fp32 [B*S*768] dQ accumulator + [B*12*S] row dot(dO, O)
      if (a <= 0 || b <= 0) break;
      return 4 * (a * b * kHidden + a * 12 * b);
    case STK_WS_LINEAR_CE_FWD: { // a = rows, b = vocabulary: fp32 (max, sumexp) per 128-column slab + target logits
      if (a <= 0 || b <= 0) break;
      const int64_t pitch = 2 * ((b + 255) / 256);
      return round_up(4 * (a * pitch * 2 + a), 256);
    }
    case STK_WS_LINEAR_CE_BWD:   // a = rows, b = vocabulary: bf16 dlogit of one vocabulary chunk
      if (a <= 0 || b <= 0) break;
      return round_up(2 * a * ce_chunk_cols(static_cast<int>(a), static_cast<int>(b)), 256);
    default:
      set_error("stk_query_workspace: unknown op %d", op);
      return STK_ERR_BAD_ARG;
  }
  set_error("stk_query_workspace: sizes must be positive (op %d, %lld, %lld)", op, (long long)a, (long long)b);
  return STK_ERR_BAD_ARG;
}

extern "C" int stk_compact_labels(int device, void* stream, const int64_t* labels, int B, int width, int row_pitch,
                                  int col_offset, int vocab, int capacity, int32_t* rows_out, int32_t* labels_out,
                                  int32_t* count_out, int* err_flag) {
  STK_REQUIRE(labels && rows_out && labels_out && count_out && B > 0 && width > 0 && capacity > 0 && vocab > 0,
              "stk_compact_labels: bad arguments");
  STK_REQUIRE(static_cast<int64_t>(B) * width < (1ll << 31) && static_cast<int64_t>(B) * row_pitch < (1ll << 31),
              "stk_compact_labels: batch too large for 32-bit row indices");
  STK_CHECK_CUDA(cudaSetDevice(device));
  compact_labels_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(labels, B, width, row_pitch, col_offset, vocab,
                                                                           capacity, rows_out, labels_out, count_out,
                                                                           err_flag);
  STK_COUNT_LAUNCH();
  return STK_OK;
}

extern "C" int stk_linear_ce_fwd(int device, void* stream, const void* t_bf16, const void* w_bf16, int R, int V,
                                 const int32_t* labels, void* workspace, int64_t workspace_bytes, float* lse,
                                 float* row_loss, float* loss_count) {
  STK_REQUIRE(t_bf16 && w_bf16 && labels && workspace && lse && R > 0 && V > 0, "stk_linear_ce_fwd: bad arguments");
  const int64_t need = stk_query_workspace(STK_WS_LINEAR_CE_FWD, R, V);
  STK_REQUIRE(workspace_bytes >= need, "stk_linear_ce_fwd: workspace too small (%lld < %lld bytes)",
              (long long)workspace_bytes, (long long)need);
  STK_REQUIRE(loss_count == nullptr || row_loss != nullptr, "stk_linear_ce_fwd: loss_count needs row_loss");
  STK_CHECK_CUDA(cudaSetDevice(device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pitch = 2 * ((static_cast<int64_t>(V) + 255) / 256);
  float* part = static_cast<float*>(workspace);
  float* tgt = part + static_cast<int64_t>(R) * pitch * 2;
  // NaN: a label that matches no column (>= V; stk_compact_labels flags those) must not read as a finite loss
  STK_CHECK_CUDA(cudaMemsetAsync(tgt, 0xFF, sizeof(float) * R, s));
  StkGemmEpilogue e = {};
  e.labels = labels;
  e.ce_partial = part;
  e.ce_pitch = pitch;
  e.tgt_logit = tgt;
  int rc = stk_gemm(device, stream, 0, 0, t_bf16, kHidden, w_bf16, kHidden, R, V, kHidden, STK_EPI_CE_STATS, nullptr, 0, &e, 1);
  if (rc) return rc;
  ce_rows_kernel<<<(R + 7) / 8, 256, 0, s>>>(reinterpret_cast<const float2*>(part), pitch, tgt, labels, R, lse, row_loss);
  STK_COUNT_LAUNCH();
  if (loss_count) {
    ce_mean_kernel<<<1, 1024, 0, s>>>(row_loss, labels, R, loss_count);
    STK_COUNT_LAUNCH();
  }
  return STK_OK;
}

extern "C" int stk_linear_ce_bwd(int device, void* stream, const void* t_bf16, const void* w_bf16, int R, int V,
                                 const int32_t* labels, const float* lse, const float* scale_dev, void* workspace,
                                 int64_t workspace_bytes, float* dT, float* dW) {
  STK_REQUIRE(t_bf16 && w_bf16 && labels && lse && scale_dev && workspace && dT && dW && R > 0 && V > 0,
              "stk_linear_ce_bwd: bad arguments");
  const int64_t need = stk_query_workspace(STK_WS_LINEAR_CE_BWD, R, V);
  STK_REQUIRE(workspace_bytes >= need, "stk_linear_ce_bwd: workspace too small (%lld < %lld bytes)",
              (long long)workspace_bytes, (long long)need);
  const int C = ce_chunk_cols(R, V);
  const __nv_bfloat16* w = static_cast<const __nv_bfloat16*>(w_bf16);
  for (int c0 = 0; c0 < V; c0 += C) {
    const int n = V - c0 < C ? V - c0 : C;
    const __nv_bfloat16* wc = w + static_cast<int64_t>(c0) * kHidden;
    StkGemmEpilogue e = {};
    e.labels = labels;
    e.lse = lse;
    e.scale_dev = scale_dev;
    e.n_offset = c0;
    // dlogit[R, n] = (softmax(T W_chunk^T) - onehot) * scale            (bf16, chunk-sized workspace, row pitch C)
    int rc = stk_gemm(device, stream, 0, 0, t_bf16, kHidden, wc, kHidden, R, n, kHidden, STK_EPI_CE_DLOGIT, workspace, C, &e, 1);
    if (rc) return rc;
    // dT[R, 768] += dlogit W_chunk                                       (B = W_chunk stored [n][768]: MN-major)
    rc = stk_gemm(device, stream, 0, 1, workspace, C, wc, kHidden, R, kHidden, n, STK_EPI_F32_ADD, dT, kHidden, nullptr, 0);
    if (rc) return rc;
    // dW_chunk[n, 768] += dlogit^T T                                     (both operands MN-major, read in place)
    rc = stk_gemm(device, stream, 1, 1, workspace, C, t_bf16, kHidden, n, kHidden, R, STK_EPI_F32_ADD,
                  dW + static_cast<int64_t>(c0) * kHidden, kHidden, nullptr, 0);
    if (rc) return rc;
  }
  return STK_OK;
}
