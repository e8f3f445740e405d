import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import ops
B,S=128,512
qkv=(torch.randn(B*S,2304,device="cuda")).bfloat16()
out=torch.empty(B*S,768,dtype=torch.bfloat16,device="cuda")
for _ in range(3): ops.attention(qkv,None,B,S,out=out)
torch.cuda.synchronize()
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.attention(qkv,None,B,S,out=out)
e1.record(); torch.cuda.synchronize()
print("STK_ATTN_DEBUG", os.environ.get("STK_ATTN_DEBUG","0"), "ms", e0.elapsed_time(e1)/10)

if int(os.environ.get("STK_ATTN_DEBUG", "0")) & 64:
    import ctypes
    from stonkgs_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * 2048)()
    lib.stk_debug_attn_timeline(buf, 2048)
    t0 = buf[0]
    names = ["m:loop", "m:sread", "m:S_issued", "m:p", "m:v", "m:PV_issued", "m:pv_done"] + [""] + ["s:start", "s:bar_s", "s:max_done", "s:p_done", "e:begin", "e:pv", "e:stored", "i:setup"]
    for g in range(12):
        row = [buf[g * 16 + i] - t0 for i in range(16)]
        print(g, " ".join(f"{n}={v}" for n, v in zip(names, row) if n))
    print("S issue start (after K ready):", [buf[1024 + i] - t0 for i in range(13)])
