"""A/B of the LayerNorm-fused GEMM geometry: CTA pairs (6-CTA clusters, 256-row tiles) vs single CTAs (3-CTA clusters,
128-row tiles; STK_GEMM_PAIR=0).  Prints ms / TFLOP/s for Wo (K=768) and FFN2 (K=3072) at M = 131 072."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import ops  # noqa: E402


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


M = 131072
torch.manual_seed(0)
r = torch.randn(M, 768, device="cuda").bfloat16()
g = torch.ones(768, device="cuda")
b = torch.zeros(768, device="cuda")
for K in (768, 3072):
    x = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(768, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(768, device="cuda") * 0.1
    ms = timed(lambda: ops.linear_resid_ln(x, w, bias, r, g, b))
    print(f"PAIR={os.environ.get('STK_GEMM_PAIR', '1')} LN-fused K={K}: {ms:.3f} ms  {2.0 * M * 768 * K / ms / 1e9:.0f} TFLOP/s")
    ms = timed(lambda: ops.linear(x, w, bias, ops.EPI_BIAS_RESID, resid=r))
    print(f"PAIR={os.environ.get('STK_GEMM_PAIR', '1')} bias+resid K={K}: {ms:.3f} ms  {2.0 * M * 768 * K / ms / 1e9:.0f} TFLOP/s")
