// stk_attn_bwd.cu — attention backward (placeholder until the tcgen05 kernel lands).
#include "stk_host.h"

extern "C" int stk_attn_bwd(int, void*, const void*, const float*, int, int, const void*, const void*, const float*,
                            float*, void*) {
  stk::set_error("stk_attn_bwd: not implemented yet");
  return STK_ERR_UNSUPPORTED;
}
