"""Shared helpers of the test-suite (golden fixtures, seeded models)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BATCH_KEYS = ("input_ids", "attention_mask", "token_type_ids", "masked_lm_labels", "ent_masked_lm_labels",
              "next_sentence_labels")


def load_fixture(name):
    fix = np.load(os.path.join(GOLDEN, name + ".npz"))
    L, B, n_kg, seed_w, seed_b, full = [int(v) for v in fix["meta"]]
    batch = {k: torch.from_numpy(fix[k]) for k in BATCH_KEYS}
    return fix, dict(layers=L, batch=B, n_kg=n_kg, seed_w=seed_w, seed_b=seed_b, full_mask=bool(full)), batch


def seeded_weights(meta):
    from oracle import weights
    return weights.make_state_dict(meta["n_kg"], meta["layers"], meta["seed_w"]), \
        weights.make_kg_table(meta["n_kg"], meta["seed_w"])


def build_model(meta, sd, rows, device=None):
    from transformers import BertConfig
    from stonkgs_b200.model import STonKGsForPreTraining
    model = STonKGsForPreTraining(None, BertConfig(vocab_size=28996, num_hidden_layers=meta["layers"]), rows)
    model.load_state_dict(sd, strict=True)
    model.eval()
    if device is not None:
        model.to(device)
    return model


def grad_sample(g: torch.Tensor):
    flat = g.reshape(-1)
    n = min(64, flat.numel())
    idx = (torch.arange(n, dtype=torch.int64) * (flat.numel() - 1)) // max(n - 1, 1)
    return flat[idx]
