"""TEST INFRASTRUCTURE — generate tests/golden/*.npz from the REAL reference module.

Run in the dev container (needs /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

For each case the reference's own ``STonKGsForPreTraining`` (imported through ``ref_shim``) is
loaded with the seeded synthetic checkpoint of ``oracle.weights`` and run in ``eval()`` on the
seeded synthetic batch of ``stonkgs_b200.synthetic``; inputs and (sub-sampled) outputs are stored.
Weights are NOT stored: they are regenerated from the seed wherever the fixture is consumed.
The script also prints how far the restatement (``stonkgs_oracle``) is from the reference.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim, stonkgs_oracle as orc, weights  # noqa: E402
from stonkgs_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (num_layers, batch, n_kg, seed_weights, seed_batch, full_mask)
CASES = {
    "L2_B2_N997": (2, 2, 997, 0, 1, False),
    "L12_B2_N997": (12, 2, 997, 0, 1, False),
    "L2_B3_N3001_fullmask": (2, 3, 3001, 3, 5, True),
}
ROWS = [0, 1, 37, 255, 256, 300, 511]  # sequence positions kept in the sub-sampled fixtures
# TransE variant (transestonkgs_model.py): 256 text + 4 KG tokens, 260 positions; same tuple layout
TRANSE_CASES = {
    "transe_L2_B3_N499": (2, 3, 499, 6, 8, False),
}
TRANSE_ROWS = [0, 1, 37, 255, 256, 257, 259]


def sample_grad(g: torch.Tensor):
    flat = g.reshape(-1)
    n = min(64, flat.numel())
    idx = (torch.arange(n, dtype=torch.int64) * (flat.numel() - 1)) // max(n - 1, 1)
    return flat[idx].numpy().copy(), idx.numpy().copy()


def run_case(name, num_layers, batch_size, n_kg, seed_w, seed_b, full_mask, transe=False):
    torch.manual_seed(0)
    ROWS = TRANSE_ROWS if transe else globals()["ROWS"]  # noqa: N806
    sd = weights.make_state_dict(n_kg, num_layers, seed_w, joint_max_pos=260 if transe else 512)
    rows = weights.make_kg_table(n_kg, seed_w)
    batch = synthetic.make_batch(batch_size, n_kg, seed_b, full_mask=full_mask, kg_len=4 if transe else 256)
    # put a few labels on padded text positions and on [CLS] (legal per the input contract)
    batch["masked_lm_labels"][0, 255] = 7
    batch["masked_lm_labels"][0, 0] = 11
    if transe:   # one pair without any entity label, one with two
        batch["ent_masked_lm_labels"][0, :] = -100
        batch["ent_masked_lm_labels"][1, 2:] = torch.tensor([5, 17])

    ref = (ref_shim.load_reference_transe if transe else ref_shim.load_reference)(sd, rows, num_layers)
    t0 = time.time()
    for p in ref.parameters():
        p.grad = None
    out = ref(**batch, return_dict=True)
    out.loss.backward()
    t_ref = time.time() - t0

    fix = {k: v.numpy() for k, v in batch.items()}
    fix["meta"] = np.array([num_layers, batch_size, n_kg, seed_w, seed_b, int(full_mask)], dtype=np.int64)
    fix["rows"] = np.array(ROWS, dtype=np.int64)
    fix["loss"] = out.loss.detach().numpy()
    fix["pooler_output"] = out.pooler_output.detach().numpy()
    fix["seq_relationship_logits"] = out.seq_relationship_logits.detach().numpy()
    fix["sequence_output_rows"] = out.hidden_states.detach()[:, ROWS].numpy()
    text_logits, ent_logits = out.prediction_logits
    sel_t = batch["masked_lm_labels"].reshape(-1) != -100
    sel_e = batch["ent_masked_lm_labels"].reshape(-1) != -100
    tl = text_logits.detach().reshape(-1, text_logits.shape[-1])[sel_t]
    el = ent_logits.detach().reshape(-1, ent_logits.shape[-1])[sel_e]
    fix["text_lse"] = torch.logsumexp(tl, -1).numpy()
    fix["entity_lse"] = torch.logsumexp(el, -1).numpy()
    fix["text_logits_head"] = tl[:, :32].numpy()       # first 32 vocabulary columns of labelled rows
    fix["entity_logits_head"] = el[:, :32].numpy()
    fix["mlm_loss"] = torch.nn.functional.cross_entropy(tl, batch["masked_lm_labels"].reshape(-1)[sel_t]).numpy()
    fix["elm_loss"] = torch.nn.functional.cross_entropy(el, batch["ent_masked_lm_labels"].reshape(-1)[sel_e]).numpy()
    # the reference's KG dict at a few ids around the index quirk (stonkgs_model.py:123-141)
    probe_ids = [0, 1, 99, 100, 101, 102, 103, 104, 105, n_kg - 1, n_kg, n_kg + 1, n_kg + 2]
    fix["kg_probe_ids"] = np.array(probe_ids, dtype=np.int64)
    fix["kg_probe_rows"] = np.stack([ref.kg_backbone[i].detach().to(torch.float32).numpy() for i in probe_ids])
    # gradients: norm + 64 strided samples per live tensor; dead tensors must have grad None
    names, norms, samples = [], [], []
    dead = []
    for k, p in ref.named_parameters():
        if not p.requires_grad:
            continue
        if p.grad is None:
            dead.append(k)
            continue
        s, _ = sample_grad(p.grad)
        names.append(k)
        norms.append(float(p.grad.norm()))
        samples.append(np.pad(s, (0, 64 - len(s))))
    fix["grad_names"] = np.array(names)
    fix["grad_norms"] = np.array(norms, dtype=np.float64)
    fix["grad_samples"] = np.stack(samples).astype(np.float32)
    fix["dead_names"] = np.array(dead)

    # ---- pin the restatement against the reference --------------------------------------------
    table = orc.build_kg_table(sd, rows)
    o, grads = orc.forward_backward(sd, table, batch, text_len=256)
    d_pool = (o["pooler_output"].detach() - out.pooler_output.detach()).abs().max().item()
    d_seq = (o["sequence_output"].detach() - out.hidden_states.detach()).abs().max().item()
    d_loss = abs(o["loss"].item() - out.loss.item())
    worst = 0.0
    ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    assert sorted(ref_grads) == sorted(grads), set(ref_grads) ^ set(grads)
    for k, g in grads.items():
        denom = ref_grads[k].abs().max().item() + 1e-30
        if "attention.self.key.bias" in k:  # analytically zero on both sides
            continue
        worst = max(worst, (g - ref_grads[k]).abs().max().item() / denom)
    tab_ok = all(torch.equal(table[i], ref.kg_backbone[i].to(torch.float32)) for i in probe_ids)
    print(f"[{name}] ref fwd+bwd {t_ref:.1f}s | oracle vs reference: pooler max|d|={d_pool:.3g} "
          f"seq max|d|={d_seq:.3g} |dloss|={d_loss:.3g} worst rel grad diff={worst:.3g} "
          f"kg table rows bit-equal={tab_ok} | dead={len(dead)} live={len(names)}")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fix)


# fine-tuning model: name -> (num_layers, batch, n_kg, seed_weights, seed_batch, num_labels)
CLS_CASES = {
    "cls_L2_B3_N997_K5": (2, 3, 997, 4, 9, 5),
}


def classifier_state(sd, num_labels, seed):
    g = torch.Generator().manual_seed(1000 + seed)
    sd = dict(sd)
    sd["classifier.weight"] = torch.randn(num_labels, 768, generator=g) * 0.05
    sd["classifier.bias"] = torch.randn(num_labels, generator=g) * 0.1
    return sd


def run_cls_case(name, num_layers, batch_size, n_kg, seed_w, seed_b, num_labels):
    """The reference's own STonKGsForSequenceClassification (stonkgs_finetuning.py:237-346), eval(), fwd + bwd."""
    sd = classifier_state(weights.make_state_dict(n_kg, num_layers, seed_w), num_labels, seed_w)
    rows = weights.make_kg_table(n_kg, seed_w)
    b = synthetic.make_batch(batch_size, n_kg, seed_b, with_labels=False)
    labels = torch.randint(0, num_labels, (batch_size,), generator=torch.Generator().manual_seed(seed_b))
    ref = ref_shim.load_reference_classifier(sd, rows, num_layers, num_labels)
    for p in ref.parameters():
        p.grad = None
    out = ref(**b, labels=labels, return_dict=True)
    out.loss.backward()
    fix = {k: v.numpy() for k, v in b.items()}
    fix["labels"] = labels.numpy()
    fix["meta"] = np.array([num_layers, batch_size, n_kg, seed_w, seed_b, num_labels], dtype=np.int64)
    fix["loss"] = out.loss.detach().numpy()
    fix["logits"] = out.logits.detach().numpy()
    names, norms, samples, dead = [], [], [], []
    for k, p in ref.named_parameters():
        if not p.requires_grad:
            continue
        if p.grad is None:
            dead.append(k)
            continue
        s, _ = sample_grad(p.grad)
        names.append(k)
        norms.append(float(p.grad.norm()))
        samples.append(np.pad(s, (0, 64 - len(s))))
    fix["grad_names"] = np.array(names)
    fix["grad_norms"] = np.array(norms, dtype=np.float64)
    fix["grad_samples"] = np.stack(samples).astype(np.float32)
    fix["dead_names"] = np.array(dead)
    table = orc.build_kg_table(sd, rows)
    o, grads = orc.forward_backward_classifier(sd, table, dict(b, labels=labels))
    ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    assert sorted(ref_grads) == sorted(grads), set(ref_grads) ^ set(grads)
    worst = max((g - ref_grads[k]).abs().max().item() / (ref_grads[k].abs().max().item() + 1e-30)
                for k, g in grads.items() if "attention.self.key.bias" not in k)
    print(f"[{name}] oracle vs reference: logits max|d|={(o['logits'].detach() - out.logits.detach()).abs().max().item():.3g} "
          f"|dloss|={abs(o['loss'].item() - out.loss.item()):.3g} worst rel grad diff={worst:.3g} | dead={len(dead)} live={len(names)}")
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fix)


def main():
    torch.set_num_threads(os.cpu_count())
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        run_case(name, *cfg)
    for name, cfg in TRANSE_CASES.items():
        if only and name not in only:
            continue
        run_case(name, *cfg, transe=True)
    for name, cfg in CLS_CASES.items():
        if only and name not in only:
            continue
        run_cls_case(name, *cfg)


if __name__ == "__main__":
    main()
