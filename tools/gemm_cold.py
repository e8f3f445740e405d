"""Bring-up: MMA cycles per k-block of a single cold launch vs the tail of a hot burst (power-throttle probe)."""
import ctypes, os, sys, time
os.environ.setdefault("STK_GEMM_DEBUG", "1")
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import _lib, ops
lib = ctypes.CDLL(_lib.LIB_PATH)

def tl():
    buf = (ctypes.c_longlong * 4096)()
    lib.stk_debug_gemm_timeline(buf, 4096)
    return [(buf[t * 16 + 1] - buf[t * 16 + 0]) for t in range(1, 6)], [buf[(t + 1) * 16] - buf[t * 16] for t in range(1, 6)]

M, N, K = 65536, 2304, 768
a = (torch.randn(M, K, device="cuda") * 0.5).bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
torch.cuda.synchronize(); time.sleep(2.0)
ops.gemm(a, w, M=M, N=N, K=K, epilogue=ops.EPI_BIAS, out=out); torch.cuda.synchronize()
print("cold single launch: mma start->lastk per tile", *tl())
for n in (5, 20, 100, 400):
    for _ in range(n): ops.gemm(a, w, M=M, N=N, K=K, epilogue=ops.EPI_BIAS, out=out)
    torch.cuda.synchronize()
    print(f"after {n} back-to-back launches:", *tl())
import subprocess
print(subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu", "--format=csv"], capture_output=True, text=True).stdout)
