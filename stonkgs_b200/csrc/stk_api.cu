// stk_api.cu — library-level pieces of the C ABI: version, error text, launch counter, and the
// host helpers shared by the kernel translation units (tensor-map encoding, SM count).
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <string.h>

#include "stk_host.h"

namespace stk {

std::atomic<long long> g_launches{0};

static thread_local char t_error[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_error, sizeof(t_error), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return STK_ERR_CUDA;
}

int num_sms(int device) {
  static int cached[64] = {};
  int& c = cached[device & 63];
  if (c == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    c = v;
  }
  return c;
}

// SMs the persistent kernels (GEMM, attention) leave free on a device: while a data-parallel gradient bucket is in flight
// the all-reduce kernel needs a few SMs, and a persistent one-CTA-per-SM kernel whose statically scheduled CTAs do not
// all fit at launch takes about TWICE as long (the late CTAs start when the others finish): measured on 2 and 8 GPUs,
// the exposed all-reduce time was ~0.45 x the collective's duration whatever its CTA count.
static std::atomic<int> g_sm_reserve[64];

int persistent_sms(int device) {
  const int n = num_sms(device) - g_sm_reserve[device & 63].load(std::memory_order_relaxed);
  return n < 1 ? 1 : n;
}

// cuTensorMapEncodeTiled is a driver entry point; resolve it through the runtime so libstk.so does
// not link against libcuda (which does not exist on the GPU-less build box).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, int elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled not available from the CUDA driver");
    return STK_ERR_CUDA;
  }
  if (box_inner * elem_bytes != 128 || box_outer > 256) {
    set_error("make_tmap_2d: box must be 128 bytes wide and at most 256 rows (got %u x %u)", box_inner, box_outer);
    return STK_ERR_BAD_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (pitch_bytes & 15) != 0) {
    set_error("make_tmap_2d: base and pitch must be 16-byte aligned");
    return STK_ERR_BAD_ARG;
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu pitch=%llu box=%ux%u)",
              static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer,
              (unsigned long long)pitch_bytes, box_inner, box_outer);
    return STK_ERR_CUDA;
  }
  return STK_OK;
}

}  // namespace stk

extern "C" int stk_version(void) { return STK_VERSION; }

extern "C" int stk_last_error(char* buf, size_t n) {
  if (buf && n) {
    strncpy(buf, stk::t_error, n - 1);
    buf[n - 1] = 0;
  }
  return static_cast<int>(strlen(stk::t_error));
}

extern "C" int stk_set_sm_reserve(int device, int n) {
  if (n < 0 || n >= stk::num_sms(device)) {
    stk::set_error("stk_set_sm_reserve: n must be in [0, %d)", stk::num_sms(device));
    return STK_ERR_BAD_ARG;
  }
  return stk::g_sm_reserve[device & 63].exchange(n, std::memory_order_relaxed);
}

extern "C" long long stk_launch_count(void) { return stk::g_launches.load(std::memory_order_relaxed); }
