"""Epilogue-vs-shape attribution for the encoder GEMMs (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stonkgs_b200 import ops
from tools.gpu_probe import _mk, _time

M = 131072
for (N, K, name) in [(768, 768, "wo"), (3072, 768, "ffn1"), (768, 3072, "ffn2"), (2304, 768, "qkv")]:
    a = _mk(M, K, "cuda", 0.5); w = _mk(N, K, "cuda", 0.05)
    bias = torch.zeros(N, device="cuda"); r = _mk(M, N, "cuda")
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    res = {}
    for epi, en in [(ops.EPI_BIAS, "bias"), (ops.EPI_BIAS_RESID, "resid"), (ops.EPI_BIAS_GELU, "gelu")]:
        ms = _time(lambda: ops.gemm(a, w, M=M, N=N, K=K, epilogue=epi, bias=bias, resid=r if epi == ops.EPI_BIAS_RESID else None, out=out))
        res[en] = round(2.0 * M * N * K / ms / 1e9)
    ms_t = _time(lambda: torch.nn.functional.linear(a, w))
    print(name, f"M={M} N={N} K={K}", res, "cublas", round(2.0 * M * N * K / ms_t / 1e9), flush=True)
