"""Synthetic batches that follow the reference's input contract.

The contract is defined by the reference's offline pre-processing
(``data/indra_for_pretraining.py:33-77,229-239`` and ``models/stonkgs_for_embeddings.py:100-155``):

* ``input_ids``       int64 ``[B, 512]``: 256 WordPiece ids (``[CLS]`` = 101 first, ``[SEP]`` = 102
  after the text, ``[PAD]`` = 0 tail) followed by 256 KG random-walk ids (two walks of 127 nodes,
  each closed by 102; unknown nodes are 100, masked ones 103);
* ``attention_mask``  text padding mask followed by ones;
* ``token_type_ids``  zeros (text half) then ones (KG half);
* ``masked_lm_labels`` / ``ent_masked_lm_labels`` int64 ``[B, 256]`` with -100 = ignore and
  ``int(256 * 0.15) = 38`` labelled positions per half, sampled from *all* positions including
  ``[CLS]``/``[SEP]``/``[PAD]`` (``indra_for_pretraining.py:55-58``);
* ``next_sentence_labels`` int64 ``[B]`` (0 = matched pair, 1 = corrupted).

Seeds and distributions are those of SURVEY.md §8d so that every consumer (tests, bench, golden
generator) sees the same batch for the same arguments.
"""
from __future__ import annotations

import torch

VOCAB = 28996
HALF = 256
CLS_ID, SEP_ID, MASK_ID, UNK_ID, PAD_ID = 101, 102, 103, 100, 0
LABELLED_PER_HALF = int(HALF * 0.15)  # 38


def make_batch(batch: int, n_kg: int, seed: int = 1, full_mask: bool = False, with_labels: bool = True,
               kg_len: int = HALF):
    """Return a dict of CPU int64 tensors shaped like one collated reference batch.
    ``kg_len`` = 256 (default) is the STonKGs random-walk half; ``kg_len`` = 4 is the TransE variant's
    ``[source, relation, target, [SEP]]`` part (transe_indra_for_pretraining.py:113-160)."""
    g = torch.Generator().manual_seed(seed)
    text = torch.randint(0, VOCAB, (batch, HALF), generator=g)
    text[:, 0] = CLS_ID
    lengths = torch.randint(32, HALF + 1, (batch,), generator=g)
    if full_mask:
        lengths[:] = HALF
    pos = torch.arange(HALF).unsqueeze(0)
    text_mask = (pos < lengths.unsqueeze(1)).long()
    # [SEP] closes the text, [PAD] fills the rest
    text = torch.where(pos == (lengths.unsqueeze(1) - 1), torch.full_like(text, SEP_ID), text)
    text = torch.where(pos >= lengths.unsqueeze(1), torch.full_like(text, PAD_ID), text)

    kg = torch.randint(0, n_kg, (batch, kg_len), generator=g)
    if kg_len == HALF:
        # a sprinkling of [UNK]/[MASK] so the three LM-backbone rows of the KG table are exercised
        special = torch.rand((batch, HALF), generator=g)
        kg = torch.where(special < 0.02, torch.full_like(kg, UNK_ID), kg)
        kg = torch.where((special >= 0.02) & (special < 0.10), torch.full_like(kg, MASK_ID), kg)
        kg[:, 127] = SEP_ID
        kg[:, 255] = SEP_ID
    else:
        kg[:, kg_len - 1] = SEP_ID
        if batch > 1:
            kg[1, 0] = MASK_ID

    out = {
        "input_ids": torch.cat([text, kg], dim=1),
        "attention_mask": torch.cat([text_mask, torch.ones_like(kg)], dim=1),
        "token_type_ids": torch.cat([torch.zeros_like(text), torch.ones_like(kg)], dim=1),
    }
    if with_labels:
        mlm = torch.full((batch, HALF), -100, dtype=torch.long)
        elm = torch.full((batch, kg_len), -100, dtype=torch.long)
        n_ent = LABELLED_PER_HALF if kg_len == HALF else 1   # the reference labels int(0.15 * 4) = 0 entity positions;
        for b in range(batch):                               # one per pair keeps the entity loss finite
            p = torch.randperm(HALF, generator=g)[:LABELLED_PER_HALF]
            mlm[b, p] = torch.randint(0, VOCAB, (LABELLED_PER_HALF,), generator=g)
            q = torch.randperm(kg_len, generator=g)[:n_ent]
            elm[b, q] = torch.randint(0, n_kg, (n_ent,), generator=g)
        out["masked_lm_labels"] = mlm
        out["ent_masked_lm_labels"] = elm
        out["next_sentence_labels"] = torch.randint(0, 2, (batch,), generator=g)
    return out
