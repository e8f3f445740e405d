"""TEST INFRASTRUCTURE — deterministic synthetic checkpoint for the STonKGs hot path.

Produces a state dict with exactly the key layout of the reference model
(``STonKGsForPreTraining.state_dict()``; SURVEY §8b: ``bert.*`` + ``lm_backbone.*`` + ``cls.*``;
413 keys at 12 layers) from a seed, so that the reference module (dev container), the oracle
restatement and the CUDA product can all be loaded with bit-identical weights on any box that
has the same torch build — the golden fixtures only need to store inputs and outputs.

Layout follows the HF BERT modules the reference instantiates
(``stonkgs_model.py:99,103,107``; HF ``modeling_bert.py`` BertEmbeddings/BertLayer/BertPooler/
BertPreTrainingHeads) and ``STonKGsELMPredictionHead.__init__`` (``stonkgs_model.py:40-60``).
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

VOCAB = 28996  # BioBERT cased vocabulary (dmis-lab/biobert-v1.1)
HIDDEN = 768
INTER = 3072
MAX_POS = 512
TYPES = 2
HEADS = 12
LN_EPS = 1e-12


def bert_keys(prefix: str, num_layers: int, max_pos: int = MAX_POS):
    """(key, shape, kind) for one HF BertModel (with pooler)."""
    H, I = HIDDEN, INTER
    MAX_POS = max_pos  # noqa: N806
    out = [
        (f"{prefix}embeddings.word_embeddings.weight", (VOCAB, H), "w"),
        (f"{prefix}embeddings.position_embeddings.weight", (MAX_POS, H), "w"),
        (f"{prefix}embeddings.token_type_embeddings.weight", (TYPES, H), "w"),
        (f"{prefix}embeddings.LayerNorm.weight", (H,), "g"),
        (f"{prefix}embeddings.LayerNorm.bias", (H,), "b"),
    ]
    for l in range(num_layers):
        p = f"{prefix}encoder.layer.{l}."
        out += [
            (p + "attention.self.query.weight", (H, H), "w"),
            (p + "attention.self.query.bias", (H,), "b"),
            (p + "attention.self.key.weight", (H, H), "w"),
            (p + "attention.self.key.bias", (H,), "b"),
            (p + "attention.self.value.weight", (H, H), "w"),
            (p + "attention.self.value.bias", (H,), "b"),
            (p + "attention.output.dense.weight", (H, H), "w"),
            (p + "attention.output.dense.bias", (H,), "b"),
            (p + "attention.output.LayerNorm.weight", (H,), "g"),
            (p + "attention.output.LayerNorm.bias", (H,), "b"),
            (p + "intermediate.dense.weight", (I, H), "w"),
            (p + "intermediate.dense.bias", (I,), "b"),
            (p + "output.dense.weight", (H, I), "w"),
            (p + "output.dense.bias", (H,), "b"),
            (p + "output.LayerNorm.weight", (H,), "g"),
            (p + "output.LayerNorm.bias", (H,), "b"),
        ]
    out += [
        (f"{prefix}pooler.dense.weight", (H, H), "w"),
        (f"{prefix}pooler.dense.bias", (H,), "b"),
    ]
    return out


def cls_keys(n_kg: int):
    """Keys of ``cls`` = BertPreTrainingHeads with the STonKGs ELM head swapped in."""
    H = HIDDEN
    return [
        ("cls.predictions.bias", (VOCAB,), "dead"),
        ("cls.predictions.text_bias", (VOCAB,), "dead"),
        ("cls.predictions.entity_bias", (n_kg,), "dead"),
        ("cls.predictions.transform.dense.weight", (H, H), "w"),
        ("cls.predictions.transform.dense.bias", (H,), "b"),
        ("cls.predictions.transform.LayerNorm.weight", (H,), "g"),
        ("cls.predictions.transform.LayerNorm.bias", (H,), "b"),
        ("cls.predictions.decoder.weight", (VOCAB, H), "dead_w"),
        ("cls.predictions.decoder.bias", (VOCAB,), "dead"),
        ("cls.predictions.decoder.text_bias", (VOCAB,), "dead"),
        ("cls.predictions.decoder.entity_bias", (n_kg,), "dead"),
        ("cls.predictions.text_decoder.weight", (VOCAB, H), "w"),
        ("cls.predictions.entity_decoder.weight", (n_kg, H), "w"),
        ("cls.seq_relationship.weight", (2, H), "w"),
        ("cls.seq_relationship.bias", (2,), "b"),
    ]


def all_keys(n_kg: int, num_layers: int = 12, joint_max_pos: int = MAX_POS):
    """``joint_max_pos`` = 260 gives the TransE variant's checkpoint (transestonkgs_model.py:93): only
    ``bert.embeddings.position_embeddings`` changes shape, the LM backbone keeps its 512 positions."""
    return bert_keys("bert.", num_layers, joint_max_pos) + bert_keys("lm_backbone.", num_layers) + cls_keys(n_kg)


def make_state_dict(n_kg: int, num_layers: int = 12, seed: int = 0, std: float = 0.02, joint_max_pos: int = MAX_POS):
    """Seeded synthetic checkpoint.

    Weights ~ N(0, std) (the BERT initialiser range), biases ~ N(0, std) (non-zero on purpose so
    that every bias path is exercised), LayerNorm gains 1 + N(0, std).  The dead tensors of the
    reference head (never read by ``forward``: ``stonkgs_model.py:55-56`` and the inherited
    ``decoder``) are zero, except ``decoder.weight`` which gets a cheap deterministic pattern.
    Each tensor draws from its own generator seeded by (seed, index) so that changing the number
    of layers or ``n_kg`` never shifts the other tensors.
    """
    sd = OrderedDict()
    for idx, (key, shape, kind) in enumerate(all_keys(n_kg, num_layers, joint_max_pos)):
        if kind in ("dead",):
            sd[key] = torch.zeros(shape, dtype=torch.float32)
            continue
        if kind == "dead_w":
            sd[key] = torch.full(shape, 0.01, dtype=torch.float32)
            continue
        g = torch.Generator().manual_seed(seed * 100003 + idx * 7919 + 17)
        t = torch.randn(shape, generator=g, dtype=torch.float32) * std
        if kind == "g":
            t = t + 1.0
        sd[key] = t
    return sd


def make_kg_table(n_kg: int, seed: int = 0) -> np.ndarray:
    """Synthetic node2vec file content: ``n_kg`` rows of 768 float32-exact values.

    The reference parses a TSV of ``repr(float32)`` values into float64 rows
    (``kg_baseline_model.py:270-280``, ``node2vec.py:350-354``); values are therefore exactly
    representable in float32, which is what this returns.
    """
    return np.random.default_rng(seed).standard_normal((n_kg, HIDDEN)).astype(np.float32)
