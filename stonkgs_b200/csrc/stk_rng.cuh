// stk_rng.cuh — counter-based dropout decisions shared by the row kernels and the attention kernels.
//
// HF BERT applies nn.Dropout(p = 0.1) at three kinds of sites (modeling_bert.py:110 embeddings, :132 attention
// probabilities, :297 / :355 dense outputs before the residual sum); torch draws its masks from a Philox stream tied to
// its own launch geometry, which no other implementation can reproduce.  Here every decision is a pure function
//     keep(seed, site, row, col) = byte (col & 3) of lowbias32(row_key(seed, site, row) + (col >> 2) * GOLDEN) >= thr
// with thr = round(256 p) (26 for p = 0.1: drop probability 26/256, survivors scaled by 256 / (256 - thr), so the
// expectation is exact), so that forward and backward kernels regenerate identical masks without storing them and
// oracle/dropout_oracle.py restates them in numpy for parity tests with the masks injected into the fp32 reference.
// lowbias32 is the 2-multiply integer finaliser (Wellons' "lowbias32"): one hash covers four elements.
#pragma once
#include <stdint.h>

namespace stk {

__host__ __device__ __forceinline__ uint32_t lowbias32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
// one per (site, row); hoisted out of the element loops
__host__ __device__ __forceinline__ uint32_t drop_row_key(uint32_t seed, uint32_t site, uint32_t row) {
  return lowbias32(lowbias32(seed ^ (site * 0x85EBCA6Bu)) + row);
}
// 4 decision bytes for columns 4*c4 .. 4*c4 + 3
__host__ __device__ __forceinline__ uint32_t drop_bytes(uint32_t row_key, uint32_t c4) {
  return lowbias32(row_key + c4 * 0x9E3779B9u);
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t bytes, int k, uint32_t thr) {
  return ((bytes >> (8 * k)) & 0xffu) >= thr;
}
__host__ __device__ __forceinline__ float drop_scale(uint32_t thr) { return 256.0f / static_cast<float>(256u - thr); }

}  // namespace stk
