"""TEST INFRASTRUCTURE — numpy restatement of the input pipeline.

* ``assemble_pairs``: the joint-sequence assembly of the reference, stonkgs_for_embeddings.py:100-135 (token type
  ids :102, text ids / padding mask :106-113, walk lookup with UNK fallback :117-126, ``walk + [SEP] + walk +
  [SEP]`` :127, attention mask :130).
* ``mask_tokens``: ``replace_mlm_tokens`` (indra_for_pretraining.py:33-77): ``int(len * 0.15)`` distinct positions
  out of ALL positions (:55-58), 80 % ``[MASK]`` (:62-63), 10 % unchanged (:66-67), 10 % random id (:69-70), labels =
  original ids (:75), -100 elsewhere (:52).  The reference draws from Python's ``random``; this restatement and the
  CUDA kernel (csrc/stk_inputs.cu) draw from Philox4x32-10 with counter (position, row, half, step) and key = seed,
  so the two can be compared bit for bit.  Parity with the reference's own stream is distributional only (header of
  stk_inputs.cu).  Only tests/ may import this module.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (Salmon et al., SC'11): uint32 arrays in, four uint32 arrays out."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = M0 * c0.astype(np.uint64)
            p1 = M1 * c2.astype(np.uint64)
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & MASK32).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & MASK32).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32(k0 + W0)
            k1 = np.uint32(k1 + W1)
    return c0, c1, c2, c3


def assemble_pairs(text_ids, text_mask, src_node, tgt_node, walks, unk_id=100, sep_id=102):
    """int arrays [n,256], [n,256] or None, [n], [n], [num_nodes, walk_len] -> three int64 arrays [n, 512]."""
    n = text_ids.shape[0]
    num_nodes, walk_len = walks.shape
    half = 256
    kg = np.full((n, half), sep_id, dtype=np.int64)
    for which, nodes in enumerate((src_node, tgt_node)):
        known = (nodes >= 0) & (nodes < num_nodes)
        w = np.full((n, walk_len), unk_id, dtype=np.int64)
        w[known] = walks[nodes[known]]
        start = which * (walk_len + 1)
        kg[:, start:start + walk_len] = w
    input_ids = np.concatenate([text_ids.astype(np.int64), kg], axis=1)
    mask = np.ones((n, half), dtype=np.int64) if text_mask is None else text_mask.astype(np.int64)
    attention_mask = np.concatenate([mask, np.ones((n, half), dtype=np.int64)], axis=1)
    token_type_ids = np.concatenate([np.zeros((n, half), dtype=np.int64), np.ones((n, half), dtype=np.int64)], axis=1)
    return input_ids, attention_mask, token_type_ids


def mask_tokens(input_ids, vocab_len, kg_vocab_len, mask_id=103, n_pick=38, seed=0, step=0, first_row=0):
    """Returns (masked input_ids int64 [n,512], mlm_labels [n,256], elm_labels [n,256])."""
    ids = np.array(input_ids, dtype=np.int64, copy=True)
    n = ids.shape[0]
    labels = []
    pos = np.arange(256, dtype=np.uint32)[None, :]
    rows = (np.arange(n, dtype=np.int64) + first_row).astype(np.uint32)[:, None]
    for half, vocab in ((0, vocab_len), (1, kg_vocab_len)):
        r0, r1, r2, r3 = philox4x32_10(pos, rows, np.uint32(half), np.uint32(step), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
        # rank of every key in its row (ties: lower position first) = place in a random permutation
        order = np.argsort(r0, axis=1, kind="stable")
        rank = np.empty_like(order)
        np.put_along_axis(rank, order, np.broadcast_to(np.arange(256), order.shape), axis=1)
        picked = rank < n_pick
        sl = slice(half * 256, (half + 1) * 256)
        orig = ids[:, sl].copy()
        repl = np.where(r1 < np.uint32(3435973836), np.int64(mask_id),
                        np.where(r2 < np.uint32(2147483648), orig, (r3 % np.uint32(vocab)).astype(np.int64)))
        ids[:, sl] = np.where(picked, repl, orig)
        labels.append(np.where(picked, orig, np.int64(-100)))
    return ids, labels[0], labels[1]
