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
