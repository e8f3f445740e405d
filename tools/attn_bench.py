"""Attention micro-benchmark (bring-up tool): forward / backward, with and without key bias and dropout.

    python tools/attn_bench.py [B]

Prints ms and TFLOP/s per case (CUDA events, 3 warm-up + 10 timed launches, inputs far larger than L2).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stonkgs_b200 import ops  # noqa: E402


def timed(fn, n=10):
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


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    torch.manual_seed(0)
    for S, biased in ((512, False), (512, True), (256, False)):
        qkv = (torch.randn(B * S, 2304, device="cuda") * 0.8).bfloat16()
        bias = None
        if biased:   # the bench's shape: text half padded after a random length in [32, 256], KG half full
            mask = torch.ones(B, S, dtype=torch.int64, device="cuda")
            lens = torch.randint(32, 257, (B,), device="cuda")
            mask[:, :256] = (torch.arange(256, device="cuda")[None, :] < lens[:, None]).long()
            bias = ops.mask_to_bias(mask)
        out = torch.empty(B * S, 768, dtype=torch.bfloat16, device="cuda")
        fl = 4.0 * B * 12 * S * S * 64
        ms = timed(lambda: ops.attention(qkv, bias, B, S, out=out))
        print(f"fwd  S={S} biased={int(biased)} B={B}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
        d = ops.Drop(seed=1, site=3, p=0.1)
        ms = timed(lambda: ops.attention(qkv, bias, B, S, out=out, drop=d))
        print(f"fwd+dropout S={S} biased={int(biased)}: {ms:.3f} ms  {fl / ms / 1e9:.0f} TFLOP/s")
        if S == 512:
            o, lse = ops.attention(qkv, bias, B, S, save_lse=True)
            dout = (torch.randn(B * S, 768, device="cuda") * 0.5).bfloat16()
            ms = timed(lambda: ops.attention_bwd(qkv, bias, B, S, o, dout, lse))
            print(f"bwd  S={S} biased={int(biased)}: {ms:.3f} ms  {2.5 * fl / ms / 1e9:.0f} TFLOP/s")
            o, lse = ops.attention(qkv, bias, B, S, save_lse=True, drop=d)
            ms = timed(lambda: ops.attention_bwd(qkv, bias, B, S, o, dout, lse, drop=d))
            print(f"bwd+dropout S={S} biased={int(biased)}: {ms:.3f} ms  {2.5 * fl / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
