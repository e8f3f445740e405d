"""``get_stonkgs_embeddings`` — batched, sharded embedding extraction.

Reference: ``src/stonkgs/models/stonkgs_for_embeddings.py:158-186``.  The reference loops over the
pre-processed rows with batch size 1, autograd on, passing the label columns so that the loss and
the dense logits are computed and thrown away, and appends ``pooler_output[0].tolist()`` to a
DataFrame (``DataFrame.append``, removed in pandas 2).  Here the same contract — a DataFrame with
``input_ids`` / ``attention_mask`` / ``token_type_ids`` columns in, a DataFrame with one
``embedding`` column of 768-float lists out, same row order — runs as batched ``no_grad`` forwards
of the CUDA path with the heads skipped; with ``torch.distributed`` initialised every rank embeds a
contiguous shard of the rows and rank 0 gathers the result (no data-path collective).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch

from .model import STonKGsForPreTraining


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n rows for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def embed_arrays(model: STonKGsForPreTraining, input_ids: np.ndarray, attention_mask: Optional[np.ndarray],
                 token_type_ids: Optional[np.ndarray], batch_size: int = 256) -> np.ndarray:
    """Pooled 768-d embeddings for int64 host arrays [n, 512]; pinned staging + async copies."""
    n = input_ids.shape[0]
    dev = model.bert.pooler.dense.weight.device
    out = torch.empty((n, 768), dtype=torch.float32, pin_memory=True)

    def stage(a, lo, hi):
        if a is None:
            return None
        t = torch.from_numpy(np.ascontiguousarray(a[lo:hi], dtype=np.int64)).pin_memory()
        return t.to(dev, non_blocking=True)

    for lo in range(0, n, batch_size):
        hi = min(lo + batch_size, n)
        model._check_ids(torch.from_numpy(np.ascontiguousarray(input_ids[lo:hi], dtype=np.int64)))
        pooled = model.embed(stage(input_ids, lo, hi), stage(attention_mask, lo, hi), stage(token_type_ids, lo, hi))
        out[lo:hi].copy_(pooled, non_blocking=True)
    torch.cuda.synchronize(dev)
    model._raise_on_bad_ids()
    return out.numpy()


def get_stonkgs_embeddings(preprocessed_df: pd.DataFrame, pretrained_stonkgs_model_name: Optional[str] = None,
                           list_of_indices: Optional[List] = None, *, model: Optional[STonKGsForPreTraining] = None,
                           batch_size: int = 256, _embed_fn=None) -> pd.DataFrame:
    """Reference signature (stonkgs_for_embeddings.py:158-162) plus two keyword-only extras:
    an already constructed ``model`` and the ``batch_size`` (``_embed_fn`` lets the CPU test-suite
    exercise the sharding / gathering logic without a GPU)."""
    if model is None and _embed_fn is None:
        if pretrained_stonkgs_model_name is not None:
            model = STonKGsForPreTraining.from_pretrained(pretrained_stonkgs_model_name)
        else:
            model = STonKGsForPreTraining.from_default_pretrained()
        model = model.to("cuda").eval()
    indices = list(list_of_indices) if list_of_indices is not None else list(preprocessed_df.index)
    rows = preprocessed_df.loc[indices]

    def col(name):
        return np.asarray(rows[name].tolist(), dtype=np.int64) if name in rows.columns else None

    ids, mask, types = col("input_ids"), col("attention_mask"), col("token_type_ids")
    world = torch.distributed.get_world_size() if torch.distributed.is_available() and torch.distributed.is_initialized() else 1
    rank = torch.distributed.get_rank() if world > 1 else 0
    lo, hi = shard_bounds(len(indices), rank, world)
    embed = _embed_fn if _embed_fn is not None else (lambda *a: embed_arrays(model, *a, batch_size))
    local = embed(ids[lo:hi], None if mask is None else mask[lo:hi], None if types is None else types[lo:hi])
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, local)
        local = np.concatenate(gathered, axis=0)
    return pd.DataFrame({"embedding": [r.tolist() for r in local]}, index=indices)
