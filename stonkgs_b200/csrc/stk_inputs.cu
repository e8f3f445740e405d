// stk_inputs.cu — GPU input pipeline for pre-tokenised text-triple pairs.
//
// Replaces the per-row Python of the reference's pre-processing for rows whose text is already tokenised:
//   * stk_assemble_pairs  — stonkgs_for_embeddings.py:100-135 / indra_for_pretraining.py:229-239: the 512-token
//     joint sequence  [text ids (256) | walk(source) (127) SEP walk(target) (127) SEP],  attention mask
//     [text padding mask | ones], token types [0 x 256 | 1 x 256]; unknown nodes become UNK walks;
//   * stk_mask_tokens     — replace_mlm_tokens (indra_for_pretraining.py:33-77) on both halves: exactly
//     int(256 * 0.15) = 38 distinct positions per half (all positions are candidates, including [CLS] / [SEP] /
//     [PAD], like the reference), 80 % -> [MASK], 10 % unchanged, 10 % random id; labels = original id, -100 elsewhere.
// Randomness is counter-based Philox4x32-10 keyed by (seed, step): the same stream is restated in numpy by
// oracle/inputs_oracle.py, so the kernel is checked BIT-EXACTLY (the reference's Mersenne-Twister call order is
// not reproduced; the distribution — uniform sample without replacement, 80/10/10 — is).
#include <atomic>

#include "stk_common.cuh"
#include "stk_host.h"

namespace stk {

extern std::atomic<long long> g_launches;

constexpr int kHalf = 256;

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                                              uint32_t (&out)[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0;
    c1 = lo1;
    c2 = hi0 ^ c3 ^ k1;
    c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// one block per pair, 512 threads = one per output token
__global__ void __launch_bounds__(512) assemble_pairs_kernel(const int32_t* __restrict__ text_ids,
                                                             const int32_t* __restrict__ text_mask,
                                                             const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                                             const int32_t* __restrict__ walks, int num_nodes, int walk_len,
                                                             int unk_id, int sep_id, int64_t* __restrict__ input_ids,
                                                             int64_t* __restrict__ attention_mask,
                                                             int64_t* __restrict__ token_type_ids) {
  const int64_t row = blockIdx.x;
  const int t = threadIdx.x;
  int64_t id, m, ty;
  if (t < kHalf) {
    id = __ldg(text_ids + row * kHalf + t);
    m = text_mask ? __ldg(text_mask + row * kHalf + t) : 1;
    ty = 0;
  } else {
    const int k = t - kHalf;                 // position in the KG half: walk | SEP | walk | SEP
    const int second = k > walk_len ? 1 : 0;
    const int pos = k - second * (walk_len + 1);
    if (pos == walk_len) {
      id = sep_id;
    } else if (pos > walk_len) {
      id = sep_id;                           // only reachable when 2 * (walk_len + 1) < 256: pad with SEP
    } else {
      const int node = __ldg((second ? tgt : src) + row);
      id = (node >= 0 && node < num_nodes) ? __ldg(walks + static_cast<int64_t>(node) * walk_len + pos) : unk_id;
    }
    m = 1;
    ty = 1;
  }
  input_ids[row * 512 + t] = id;
  attention_mask[row * 512 + t] = m;
  token_type_ids[row * 512 + t] = ty;
}

// one block per (pair, half), 256 threads = one per position
__global__ void __launch_bounds__(256) mask_tokens_kernel(int64_t* __restrict__ input_ids, int64_t* __restrict__ mlm_labels,
                                                          int64_t* __restrict__ elm_labels, int vocab_len, int kg_vocab_len,
                                                          int mask_id, int n_pick, uint32_t seed_lo, uint32_t seed_hi,
                                                          uint32_t step, int64_t row0) {
  __shared__ uint32_t keys[kHalf];
  const int64_t row = blockIdx.x >> 1;
  const int half = blockIdx.x & 1;
  const int t = threadIdx.x;
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(t), static_cast<uint32_t>(row0 + row), static_cast<uint32_t>(half), step, seed_lo,
                seed_hi, r);
  keys[t] = r[0];
  __syncthreads();
  // rank of this position's key (ties: lower position first) = its place in a random permutation
  int rank = 0;
  const uint32_t mine = r[0];
#pragma unroll 8
  for (int j = 0; j < kHalf; ++j) {
    const uint32_t k = keys[j];
    rank += (k < mine || (k == mine && j < t)) ? 1 : 0;
  }
  int64_t* ids = input_ids + row * 512 + half * kHalf;
  int64_t* labels = (half ? elm_labels : mlm_labels) + row * kHalf;
  const int64_t orig = ids[t];
  int64_t label = -100;
  if (rank < n_pick) {
    label = orig;
    int64_t repl;
    if (r[1] < 3435973836u) {          // < 0.8
      repl = mask_id;
    } else if (r[2] < 2147483648u) {   // < 0.5: keep the token
      repl = orig;
    } else {
      repl = r[3] % static_cast<uint32_t>(half ? kg_vocab_len : vocab_len);
    }
    ids[t] = repl;
  }
  labels[t] = label;
}

}  // namespace stk

using namespace stk;

extern "C" int stk_assemble_pairs(int device, void* stream, const int32_t* text_ids, const int32_t* text_mask,
                                  const int32_t* src_node, const int32_t* tgt_node, const int32_t* walks, int num_nodes,
                                  int walk_len, int unk_id, int sep_id, int n, int64_t* input_ids, int64_t* attention_mask,
                                  int64_t* token_type_ids) {
  STK_REQUIRE(text_ids && src_node && tgt_node && walks && input_ids && attention_mask && token_type_ids && n > 0,
              "stk_assemble_pairs: bad arguments");
  STK_REQUIRE(walk_len > 0 && 2 * (walk_len + 1) <= kHalf, "stk_assemble_pairs: two walks + two SEP must fit 256 tokens (walk_len=%d)", walk_len);
  STK_CHECK_CUDA(cudaSetDevice(device));
  assemble_pairs_kernel<<<n, 512, 0, static_cast<cudaStream_t>(stream)>>>(text_ids, text_mask, src_node, tgt_node, walks,
                                                                          num_nodes, walk_len, unk_id, sep_id, input_ids,
                                                                          attention_mask, token_type_ids);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}

extern "C" int stk_mask_tokens(int device, void* stream, int64_t* input_ids, int64_t* mlm_labels, int64_t* elm_labels, int n,
                               int vocab_len, int kg_vocab_len, int mask_id, int n_pick, uint64_t seed, uint32_t step,
                               int64_t first_row) {
  STK_REQUIRE(input_ids && mlm_labels && elm_labels && n > 0, "stk_mask_tokens: bad arguments");
  STK_REQUIRE(vocab_len > 0 && kg_vocab_len > 0 && n_pick >= 0 && n_pick <= kHalf, "stk_mask_tokens: bad sizes");
  STK_CHECK_CUDA(cudaSetDevice(device));
  mask_tokens_kernel<<<2 * n, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      input_ids, mlm_labels, elm_labels, vocab_len, kg_vocab_len, mask_id, n_pick, static_cast<uint32_t>(seed),
      static_cast<uint32_t>(seed >> 32), step, first_row);
  STK_CHECK_CUDA(cudaGetLastError());
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return STK_OK;
}
