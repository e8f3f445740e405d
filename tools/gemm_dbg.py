"""Bring-up: clock64 timeline of CTA 0 of the GEMM kernel (STK_GEMM_DEBUG=1) for the encoder shapes."""
import ctypes
import os
import sys

os.environ.setdefault("STK_GEMM_DEBUG", "1")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import _lib, ops

NAMES = ["mma:start", "mma:lastk", "e:loop", "e:tfull", "e:ld0", "e:xin0", "e:cmp0", "e:st0", "e:ld1", "e:xin1",
         "e:cmp1", "e:st1", "tma:start", "mma:firstk", "mma:starved"]


def run(name, M, N, K, epi, **kw):
    dev = "cuda"
    a = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    extra = {}
    if epi in (ops.EPI_BIAS_RESID,) or kw.get("ln"):
        extra["resid"] = torch.randn(M, N, device=dev).bfloat16()
    if kw.get("ln"):
        extra.update(ln_gamma=torch.ones(N, device=dev), ln_beta=torch.zeros(N, device=dev))
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        ops.gemm(a, w, M=M, N=N, K=K, epilogue=epi, bias=bias, out=out, **extra)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, w, M=M, N=N, K=K, epilogue=epi, bias=bias, out=out, **extra)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"== {name}: M={M} N={N} K={K} epi={epi}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.0f} TFLOP/s")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    buf = (ctypes.c_longlong * 4096)()
    lib.stk_debug_gemm_timeline(buf, 4096)
    t0 = buf[2]
    kbs = [(buf[3072 + 2 * i] - buf[3072], buf[3072 + 2 * i + 1] - buf[3072]) for i in range(min(K // 64, 16))]
    print("  tile 2 k-blocks (ready, issued):", kbs)
    if kw.get("ln"):
        names = ["mma:start", "mma:lastk", "e:loop", "e:tfull", "e:r0", "e:p1_0", "e:stats", "e:p2_0", "e:r1", "e:p1_1",
                 "io:stored", "e:p2_1", "e:published", "mma:firstk"]
        for cta in range(3):
            for t in range(2, 5):
                o = cta * 1024 + t * 16
                row = [buf[o + i] - t0 for i in range(14)]
                print(f"  cta {cta} tile {t}: " + " ".join(f"{n}={v}" for n, v in zip(names, row)) +
                      f" io:loaded={buf[o + 15] - t0} starved={buf[o + 14]}")
        return
    for t in range(1, 5):
        row = [buf[t * 16 + i] - t0 for i in range(14)] + [buf[t * 16 + 14]]
        print(f"  tile {t}: " + " ".join(f"{n}={v}" for n, v in zip(NAMES, row)))


if __name__ == "__main__":
    M = int(os.environ.get("M", 65536))
    if not os.environ.get("LN_ONLY"):
        run("Wo+resid", M, 768, 768, ops.EPI_BIAS_RESID)
        run("FFN2+resid", M, 768, 3072, ops.EPI_BIAS_RESID)
        run("QKV", M, 2304, 768, ops.EPI_BIAS)
        run("FFN1+gelu", M, 3072, 768, ops.EPI_BIAS_GELU)
    if hasattr(ops, "EPI_BIAS_RESID_LN"):
        run("Wo+resid+LN", M, 768, 768, ops.EPI_BIAS_RESID_LN, ln=True)
        run("FFN2+resid+LN", M, 768, 3072, ops.EPI_BIAS_RESID_LN, ln=True)
