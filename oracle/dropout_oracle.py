"""TEST INFRASTRUCTURE — numpy restatement of the dropout decision function (csrc/stk_rng.cuh).

The reference's dropout is ``nn.Dropout`` inside HF BERT (modeling_bert.py:110 embeddings, :132 attention
probabilities, :297 / :355 dense outputs); its masks come from torch's Philox stream and cannot be reproduced by another
implementation, so training-mode parity is defined as: the CUDA path with dropout == the fp32 oracle with THE SAME
masks injected at the same sites (``stonkgs_oracle.forward(..., drop=...)``).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np
import torch

GOLDEN = np.uint32(0x9E3779B9)


def lowbias32(x):
    x = np.asarray(x, dtype=np.uint32).copy()
    with np.errstate(over="ignore"):
        x ^= x >> np.uint32(16)
        x *= np.uint32(0x7FEB352D)
        x ^= x >> np.uint32(15)
        x *= np.uint32(0x846CA68B)
        x ^= x >> np.uint32(16)
    return x


def drop_words(row_key, c8):
    """(w0, w1) of stk_rng.cuh: w0 = lowbias32(row_key + c8 * GOLDEN); w1 = hi32 ^ lo32 of w0 * 0x9E3779B1."""
    with np.errstate(over="ignore"):
        w0 = lowbias32(np.asarray(row_key, dtype=np.uint32) + np.asarray(c8, dtype=np.uint32) * GOLDEN)
    m = w0.astype(np.uint64) * np.uint64(0x9E3779B1)
    w1 = ((m >> np.uint64(32)) ^ (m & np.uint64(0xFFFFFFFF))).astype(np.uint32)
    return w0, w1


def keep_mask(seed: int, site: int, n_rows: int, n_cols: int, thr: int) -> np.ndarray:
    """bool [n_rows, n_cols]: keep(seed, site, row, col) of stk_rng.cuh (thr = round(128 p) < 128)."""
    with np.errstate(over="ignore"):
        s = lowbias32(np.uint32(seed & 0xFFFFFFFF) ^ np.uint32((site * 0x85EBCA6B) & 0xFFFFFFFF))
        row_key = lowbias32(s + np.arange(n_rows, dtype=np.uint32))
    n8 = (n_cols + 7) // 8
    w0, w1 = drop_words(row_key[:, None], np.arange(n8, dtype=np.uint32)[None, :])          # [rows, cols/8]
    words = np.stack([w0, w1], axis=-1)                                                       # [rows, cols/8, 2]
    b = np.stack([(words >> np.uint32(8 * k)) & np.uint32(0x7F) for k in range(4)], axis=-1)  # [rows, cols/8, 2, 4]
    return (b >= np.uint32(thr)).reshape(n_rows, n8 * 8)[:, :n_cols]


class DropSpec:
    """Callable handed to the oracle: ``drop(site, tensor)`` masks and rescales ``tensor`` the way the kernels do.
    Hidden sites see [B, S, 768] (row = b*S + s); attention sites see [B, 12, S, S] (row = (b*12 + h)*S + q)."""

    def __init__(self, seed: int, p_hidden: float, p_attn: float):
        self.seed = seed & 0xFFFFFFFF
        self.thr_h = min(127, int(round(128.0 * p_hidden)))
        self.thr_a = min(127, int(round(128.0 * p_attn)))

    def __call__(self, site: int, t: torch.Tensor, attention: bool = False) -> torch.Tensor:
        thr = self.thr_a if attention else self.thr_h
        if thr == 0:
            return t
        cols = t.shape[-1]
        rows = t.numel() // cols
        keep = torch.from_numpy(keep_mask(self.seed, site, rows, cols, thr)).view(t.shape)
        return torch.where(keep, t * (128.0 / (128 - thr)), torch.zeros_like(t))


# site ids (shared with stonkgs_b200/engine.py): encoder 0 = frozen LM backbone, 1 = joint encoder
def site_embeddings(encoder: int) -> int:
    return encoder * 64 + 63


def site_attention(encoder: int, layer: int) -> int:
    return encoder * 64 + layer * 4


def site_attn_out(encoder: int, layer: int) -> int:
    return encoder * 64 + layer * 4 + 1


def site_ffn_out(encoder: int, layer: int) -> int:
    return encoder * 64 + layer * 4 + 2
