"""``get_stonkgs_embeddings`` — batched, streamed, sharded embedding extraction.

Reference: ``src/stonkgs/models/stonkgs_for_embeddings.py:158-186``.  The reference loops over the
pre-processed rows with batch size 1, autograd on, passing the label columns so that the loss and
the dense logits are computed and thrown away, and appends ``pooler_output[0].tolist()`` to a
DataFrame (``DataFrame.append``, removed in pandas 2).  Here the same contract — a DataFrame with
``input_ids`` / ``attention_mask`` / ``token_type_ids`` columns in, a DataFrame with one
``embedding`` column of 768-float lists out, same row order — runs as batched ``no_grad`` forwards
of the CUDA path with the heads skipped; with ``torch.distributed`` initialised every rank embeds a
contiguous shard of the rows and rank 0 gathers the result (no data-path collective).

Bulk extraction (BASELINE configs[4]: 10 M pairs over 2/4/8 GPUs) goes through :func:`embed_arrays`:
host id arrays in, one float32 ``[n, 768]`` NumPy array out, streamed through a small ring of reusable
pinned staging slots on a copy stream, so that the H2D copy of batch i+1 and the D2H copy of batch i-1
run under the kernels of batch i.  The DataFrame wrapper exists for API parity only: a column of Python
float lists costs ~100 x the memory of the array.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import pandas as pd
import torch

from ._lib import StkError
from .model import STonKGsForPreTraining


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n rows for this rank (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _Slot:
    """One staging slot: pinned host ids -> device ids (copy stream), pooled -> pinned host (copy stream)."""

    def __init__(self, batch: int, seq_len: int, n_cols: int, dev, out_width: int = 768):
        self.h_in = torch.empty((n_cols, batch, seq_len), dtype=torch.int64).pin_memory()
        self.h_in_np = self.h_in.numpy()
        self.d_in = torch.empty((n_cols, batch, seq_len), dtype=torch.int64, device=dev)
        self.h_out = torch.empty((batch, out_width), dtype=torch.float32).pin_memory()
        self.h_out_np = self.h_out.numpy()
        self.h2d = torch.cuda.Event()     # ids are on the device
        self.done = torch.cuda.Event()    # pooled rows are in h_out
        self.span = None                  # (lo, hi) of the batch in flight, None = free


class EmbeddingStreamer:
    """Reusable streaming state of :func:`embed_arrays` (staging ring + copy stream); ``h2d_bytes`` / ``d2h_bytes`` count
    what crossed the bus."""

    def __init__(self, model: STonKGsForPreTraining, batch_size: int = 256, slots: int = 2, columns: int = 3,
                 pooling: str = "pooler", fn=None, out_width: int = 768, cls_rows_only: bool = False,
                 skip_padding: bool = False):
        """``fn(input_ids, attention_mask, token_type_ids, err_flag=...)`` -> fp32 [m, out_width] on the device replaces
        ``model.embed`` (the fine-tuning model streams class probabilities through the same staging ring).
        ``cls_rows_only`` / ``skip_padding``: see :meth:`STonKGsForPreTraining.embed` (the plan of the padded rows is
        made from the staged host copy of the mask: no synchronisation)."""
        self.model = model
        self.pooling = pooling
        self.cls_rows_only = bool(cls_rows_only)
        self.skip_padding = bool(skip_padding) and columns >= 2
        self.fn = fn
        self.out_width = int(out_width)
        self.dev = model.bert.pooler.dense.weight.device
        if self.dev.type != "cuda":
            raise StkError("embedding extraction runs on CUDA only: move the model with .to('cuda')")
        self.batch_size = int(batch_size)
        self.seq_len = model.seq_shape.seq_len
        self.columns = columns
        self.slots = [_Slot(self.batch_size, self.seq_len, columns, self.dev, self.out_width) for _ in range(max(2, slots))]
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.err = torch.zeros(1, dtype=torch.int32, device=self.dev)   # one id-range flag for the whole stream
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _retire(self, slot: _Slot, out: np.ndarray):
        if slot.span is None:
            return
        slot.done.synchronize()
        lo, hi = slot.span
        np.copyto(out[lo:hi], slot.h_out_np[: hi - lo])
        slot.span = None

    @torch.no_grad()
    def run(self, input_ids, attention_mask=None, token_type_ids=None, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Arrays only need ``.shape[0]`` and ``a[lo:hi] -> ndarray`` (any integer dtype), so memory-mapped or tiled
        sources stream without being materialised.  ``out`` may be a caller-provided float32 [n, 768] array / memmap."""
        n = int(input_ids.shape[0])
        if out is None:
            out = np.empty((n, self.out_width), dtype=np.float32)
        cols = [input_ids, attention_mask, token_type_ids][: self.columns]
        compute = torch.cuda.current_stream(self.dev)
        model, bs = self.model, self.batch_size
        table_rows = model.kg_table.shape[0]
        for i, lo in enumerate(range(0, n, bs)):
            hi = min(lo + bs, n)
            m = hi - lo
            slot = self.slots[i % len(self.slots)]
            self._retire(slot, out)                       # the slot's previous batch (i - slots) is complete
            for c, a in enumerate(cols):
                if a is not None:
                    np.copyto(slot.h_in_np[c, :m], a[lo:hi], casting="same_kind")
            with torch.cuda.stream(self.copy_stream):
                dcols = []
                for c, a in enumerate(cols):
                    if a is None:
                        dcols.append(None)
                        continue
                    slot.d_in[c, :m].copy_(slot.h_in[c, :m], non_blocking=True)
                    dcols.append(slot.d_in[c, :m])
                    self.h2d_bytes += m * self.seq_len * 8
                slot.h2d.record(self.copy_stream)
            compute.wait_event(slot.h2d)
            if self.fn is not None:
                pooled = self.fn(*dcols, err_flag=self.err)
            else:
                pooled = model.embed(*dcols, err_flag=self.err, pooling=self.pooling, cls_rows_only=self.cls_rows_only,
                                     skip_padding=self.skip_padding and cols[1] is not None,
                                     host_mask=slot.h_in_np[1, :m] if self.skip_padding else None)
            ready = torch.cuda.Event()
            ready.record(compute)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(ready)
                pooled.record_stream(self.copy_stream)
                slot.h_out[:m].copy_(pooled, non_blocking=True)
                slot.done.record(self.copy_stream)
            self.d2h_bytes += m * self.out_width * 4
            slot.span = (lo, hi)
        for slot in self.slots:                           # drain
            self._retire(slot, out)
        if int(self.err.item()) != 0:
            self.err.zero_()
            raise KeyError(f"input id outside the text vocabulary / KG table ({table_rows} rows)")
        return out


def embed_arrays(model: STonKGsForPreTraining, input_ids, attention_mask=None, token_type_ids=None,
                 batch_size: int = 256, out: Optional[np.ndarray] = None,
                 streamer: Optional[EmbeddingStreamer] = None, pooling: str = "pooler",
                 cls_rows_only: bool = False, skip_padding: bool = False) -> np.ndarray:
    """Pooled 768-d embeddings (``pooler_output``, stonkgs_for_embeddings.py:180; ``pooling="mean"``: masked mean of the
    last hidden state) for host id arrays ``[n, 512]``.  Returns float32 ``[n, 768]``; see :class:`EmbeddingStreamer`
    for the staging scheme and :meth:`STonKGsForPreTraining.embed` for ``cls_rows_only`` / ``skip_padding``."""
    st = streamer if streamer is not None else EmbeddingStreamer(model, batch_size, pooling=pooling,
                                                                 cls_rows_only=cls_rows_only, skip_padding=skip_padding)
    return st.run(input_ids, attention_mask, token_type_ids, out=out)


def get_stonkgs_embeddings(preprocessed_df: pd.DataFrame, pretrained_stonkgs_model_name: Optional[str] = None,
                           list_of_indices: Optional[List] = None, *, model: Optional[STonKGsForPreTraining] = None,
                           batch_size: int = 256, _embed_fn=None) -> pd.DataFrame:
    """Reference signature (stonkgs_for_embeddings.py:158-162) plus two keyword-only extras:
    an already constructed ``model`` and the ``batch_size`` (``_embed_fn`` lets the CPU test-suite
    exercise the sharding / gathering logic without a GPU).

    Like the reference (:172-184) ``list_of_indices`` are POSITIONS (``iloc``), and the returned frame has a
    fresh ``RangeIndex`` (``append(..., ignore_index=True)``)."""
    if model is None and _embed_fn is None:
        if pretrained_stonkgs_model_name is not None:
            model = STonKGsForPreTraining.from_pretrained(pretrained_stonkgs_model_name)
        else:
            model = STonKGsForPreTraining.from_default_pretrained()
        model = model.to("cuda").eval()
    indices = list(list_of_indices) if list_of_indices is not None else list(range(len(preprocessed_df)))
    rows = preprocessed_df.iloc[indices]

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
    return pd.DataFrame({"embedding": [r.tolist() for r in local]})
