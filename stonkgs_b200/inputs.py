"""GPU input pipeline for pre-tokenised text-triple pairs and the binary node2vec table.

Reference: ``preprocess_df_for_embeddings_iter`` (stonkgs_for_embeddings.py:100-155) builds every row in Python —
tokenise, look the two random walks up in a dict, splice ``[SEP]``, mask 15 % of each half with ``random`` — and
``indra_to_pretraining_df`` pickles the result (indra_for_pretraining.py:292-294); the node2vec vectors live in a
TSV (node2vec.py:350-354).  Here the tokeniser stays on the host (HF fast tokenizer, third-party), everything after
it runs as two kernels over whole batches (``stk_assemble_pairs``, ``stk_mask_tokens``), and the table can be stored
as a float32 ``.npy`` + a names file that ``STonKGsForPreTraining`` memory-maps.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import check

HALF = 256
SEP_ID, MASK_ID, UNK_ID = 102, 103, 100


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _ctx(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.StkError("the input pipeline kernels need CUDA tensors (there is no CPU path)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return dev, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class WalkTable:
    """Random walks of the pre-training KG as one device int32 matrix [num_nodes, walk_len] + a name -> row map
    (the reference keeps ``{node name: [walk_len node indices]}``, stonkgs_for_embeddings.py:84-91)."""

    def __init__(self, walks_by_name: Dict[str, Sequence[int]], device="cuda"):
        self.names = list(walks_by_name)
        self.index = {n: i for i, n in enumerate(self.names)}
        mat = np.asarray([walks_by_name[n] for n in self.names], dtype=np.int32)
        if mat.ndim != 2 or 2 * (mat.shape[1] + 1) > HALF:
            raise ValueError("all walks must have the same length L with 2 * (L + 1) <= 256")
        self.walk_len = int(mat.shape[1])
        self.walks = torch.from_numpy(mat).to(device)

    def rows(self, nodes: Sequence[str]) -> np.ndarray:
        """Row index per node name, -1 for nodes that were not part of pre-training (-> UNK walk)."""
        return np.asarray([self.index.get(n, -1) for n in nodes], dtype=np.int32)


def assemble_pairs(text_ids: torch.Tensor, text_mask: Optional[torch.Tensor], src_rows: torch.Tensor,
                   tgt_rows: torch.Tensor, table: WalkTable, unk_id: int = UNK_ID, sep_id: int = SEP_ID
                   ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """int32 device tensors [n,256] / [n,256] / [n] / [n] -> int64 ``input_ids``, ``attention_mask``,
    ``token_type_ids`` [n, 512] (the batch contract of ``STonKGsForPreTraining.forward``)."""
    for t, name in ((text_ids, "text_ids"), (src_rows, "src_rows"), (tgt_rows, "tgt_rows")):
        if t.dtype != torch.int32 or not t.is_contiguous():
            raise _lib.StkError(f"{name}: expected a contiguous int32 tensor")
    n = text_ids.shape[0]
    dev, stream = _ctx(text_ids)
    ids = torch.empty((n, 512), dtype=torch.int64, device=text_ids.device)
    mask = torch.empty_like(ids)
    types = torch.empty_like(ids)
    if text_mask is not None:
        text_mask = text_mask.to(torch.int32).contiguous()
    check(_lib.load().stk_assemble_pairs(dev, stream, _ptr(text_ids), _ptr(text_mask), _ptr(src_rows), _ptr(tgt_rows),
                                         _ptr(table.walks), table.walks.shape[0], table.walk_len, unk_id, sep_id, n,
                                         _ptr(ids), _ptr(mask), _ptr(types)), "stk_assemble_pairs")
    return ids, mask, types


def mask_tokens(input_ids: torch.Tensor, vocab_len: int, kg_vocab_len: int, *, seed: int, step: int = 0,
                first_row: int = 0, mask_id: int = MASK_ID, masked_tokens_percentage: float = 0.15
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    """In-place ``replace_mlm_tokens`` on both halves of int64 ``input_ids`` [n,512]; returns
    (``masked_lm_labels``, ``ent_masked_lm_labels``) int64 [n,256].  (seed, step, first_row) select the Philox
    stream, so a data-parallel rank masks its shard exactly as a single process would."""
    if input_ids.dtype != torch.int64 or not input_ids.is_contiguous() or input_ids.shape[1] != 512:
        raise _lib.StkError("input_ids: expected a contiguous int64 tensor [n, 512]")
    n = input_ids.shape[0]
    dev, stream = _ctx(input_ids)
    mlm = torch.empty((n, HALF), dtype=torch.int64, device=input_ids.device)
    elm = torch.empty_like(mlm)
    n_pick = int(HALF * masked_tokens_percentage)     # reference :57
    check(_lib.load().stk_mask_tokens(dev, stream, _ptr(input_ids), _ptr(mlm), _ptr(elm), n, vocab_len, kg_vocab_len,
                                      mask_id, n_pick, seed & 0xFFFFFFFFFFFFFFFF, step & 0xFFFFFFFF, first_row),
          "stk_mask_tokens")
    return mlm, elm


# --------------------------------------------------------------------------------------------------
# binary node2vec table
# --------------------------------------------------------------------------------------------------
def save_kg_table(path: str, names: Sequence[str], rows: np.ndarray) -> None:
    """``<path>`` (float32 ``.npy`` [N, 768]) + ``<path>.names`` (one node name per line, file order)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    if rows.ndim != 2 or len(names) != rows.shape[0]:
        raise ValueError("names and rows disagree")
    np.save(path, rows)
    with open(path + ".names", "w") as f:
        f.write("\n".join(names) + "\n")


def load_kg_table(path: str):
    """(names, float32 [N, 768]) from the binary form written by ``save_kg_table`` (memory-mapped)."""
    rows = np.load(path, mmap_mode="r")
    names_path = path + ".names"
    if os.path.exists(names_path):
        with open(names_path) as f:
            names = [ln.rstrip("\n") for ln in f if ln.strip()]
    else:
        names = [f"n{i}" for i in range(rows.shape[0])]
    if len(names) != rows.shape[0]:
        raise ValueError(f"{names_path}: {len(names)} names for {rows.shape[0]} rows")
    return names, rows


def tsv_to_binary(tsv_path: str, out_path: str) -> None:
    """One-off conversion of the reference's ``embeddings_best_model.tsv`` (node2vec.py:350-354)."""
    from .model import prepare_df
    names, rows = prepare_df(tsv_path)
    save_kg_table(out_path, names, rows)
