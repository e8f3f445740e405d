// stk_rng.cuh — counter-based dropout decisions shared by the row kernels and the attention kernels.
//
// HF BERT applies nn.Dropout(p = 0.1) at three kinds of sites (modeling_bert.py:110 embeddings, :132 attention
// probabilities, :297 / :355 dense outputs before the residual sum); torch draws its masks from a Philox stream tied to
// its own launch geometry, which no other implementation can reproduce.  Here every decision is a pure function
//     words(seed, site, row, col >> 3) = (w0, w1):  w0 = lowbias32(row_key(seed, site, row) + (col >> 3) * GOLDEN),
//                                                   w1 = hi32 ^ lo32 of the 64-bit product w0 * 0x9E3779B1
//     keep(seed, site, row, col)       = (byte (col & 3) of (col & 4 ? w1 : w0)) & 0x7f  >=  thr
// with thr = round(128 p) (13 for p = 0.1: drop probability 13/128, survivors scaled by 128 / (128 - thr), so the
// expectation is exact), so that forward and backward kernels regenerate identical masks without storing them and
// oracle/dropout_oracle.py restates them in numpy for parity tests with the masks injected into the fp32 reference.
// lowbias32 is the 2-multiply integer finaliser (Wellons' "lowbias32"); one hash + one wide multiply cover eight
// elements, and the four 7-bit comparisons of a word are done at once (drop_signs), because the attention softmax
// warps have few integer issue slots to spare beside their exponentials.
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
// decision words of columns 8*c8 .. 8*c8 + 3 (w0) and 8*c8 + 4 .. 8*c8 + 7 (w1)
__host__ __device__ __forceinline__ void drop_words(uint32_t row_key, uint32_t c8, uint32_t& w0, uint32_t& w1) {
  w0 = lowbias32(row_key + c8 * 0x9E3779B9u);
  const uint64_t m = static_cast<uint64_t>(w0) * 0x9E3779B1ull;
  w1 = static_cast<uint32_t>(m >> 32) ^ static_cast<uint32_t>(m);
}
// thr replicated into the four bytes (thr < 128)
__host__ __device__ __forceinline__ uint32_t drop_thr4(uint32_t thr) { return thr * 0x01010101u; }
__host__ __device__ __forceinline__ uint32_t drop_thr4_of(uint32_t thr) { return drop_thr4(thr); }   // (a local may shadow drop_thr4)
// bit 7 of byte k of the result is the keep decision of column k of the word: (byte & 0x7f) + 0x80 - thr has bit 7 set
// iff (byte & 0x7f) >= thr, and no byte borrows from its neighbour
__host__ __device__ __forceinline__ uint32_t drop_signs(uint32_t w, uint32_t thr4) {
  return ((w & 0x7f7f7f7fu) | 0x80808080u) - thr4;
}
__host__ __device__ __forceinline__ bool drop_keep(uint32_t signs, int k) { return (signs >> (8 * k + 7)) & 1u; }
__host__ __device__ __forceinline__ float drop_scale(uint32_t thr) { return 128.0f / static_cast<float>(128u - thr); }

#ifdef __CUDACC__
// 0xffffffff where column k of the word is kept, else 0 (PRMT replicates the sign bit of the selected byte)
template <int K>
__device__ __forceinline__ uint32_t drop_mask32(uint32_t signs) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(signs), "n"(0x8888 + 0x1111 * K));
  return m;
}
// bf16x2 mask of the column pair (2P, 2P + 1): low half follows column 2P, high half column 2P + 1
template <int P>
__device__ __forceinline__ uint32_t drop_mask16x2(uint32_t signs) {
  uint32_t m;
  asm("prmt.b32 %0, %1, %1, %2;" : "=r"(m) : "r"(signs), "n"(P == 0 ? 0x9988 : 0xBBAA));
  return m;
}
#endif

}  // namespace stk
