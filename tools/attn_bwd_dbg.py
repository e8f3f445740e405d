import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stonkgs_b200 import ops, _lib
B, S = 64, 512
qkv = torch.randn(B * S, 2304, device="cuda").bfloat16()
out, lse = ops.attention(qkv, None, B, S, save_lse=True)
dout = torch.randn(B * S, 768, device="cuda").bfloat16()
for _ in range(3): ops.attention_bwd(qkv, None, B, S, out, dout, lse)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention_bwd(qkv, None, B, S, out, dout, lse)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print("attn_bwd B=64 S=512 ms", ms, "TFLOP/s", 10.0 * B * 12 * S * S * 64 / ms / 1e9)
if int(os.environ.get("STK_ATTN_DEBUG", "0")) & 64:
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * 256)()
    lib.stk_debug_attn_bwd_timeline(buf, 256)
    t0 = buf[0]
    names = {1: "m:loop", 2: "m:p_ready", 3: "m:issued", 8: "c:start", 9: "c:S_ready", 10: "c:PdS_written", 12: "c:prev_dq_drained"}
    for i in range(4):
        print(i, " ".join(f"{n}={buf[i * 16 + k] - t0}" for k, n in names.items()))
    print("cta: entry=%d prologue_done=%d loop_done=%d dvdk_done=%d stored=%d exit=%d" % tuple(buf[64 + k] - t0 for k in range(6)))
    for g in range(4):
        print("PdS_written per warp 1..7, iteration", g, [buf[128 + g * 8 + w] - t0 for w in range(1, 8)])
