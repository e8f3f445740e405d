// stk_host.h — host-side helpers shared by the C-ABI translation units.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/stk.h"

namespace stk {

// Thread-local last-error text (retrievable through stk_last_error).
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int num_sms(int device);
int persistent_sms(int device);   // num_sms minus the SMs reserved for a concurrent collective (stk_set_sm_reserve)

#define STK_CHECK_CUDA(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::stk::cuda_fail(_e, #expr); \
  } while (0)

#define STK_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ::stk::set_error(__VA_ARGS__);  \
      return STK_ERR_BAD_ARG;         \
    }                                 \
  } while (0)

// 2-D tiled tensor map with 128-byte swizzle.  `inner`/`outer` are the global extents in elements,
// `pitch_bytes` the byte stride between outer rows, `box_inner * elem_bytes` must be 128.
// Out-of-bounds parts of a box read as zero and are clipped on store.
int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer);

}  // namespace stk
