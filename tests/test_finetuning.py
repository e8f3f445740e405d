"""Fine-tuning model: oracle restatement vs the golden fixture generated from the reference's own
``STonKGsForSequenceClassification`` (CPU), and the CUDA drop-in vs the oracle (GPU).

Tolerances (bf16 tensor-core compute vs fp32 reference): logits atol 4e-2, loss rtol 5e-3, gradients cosine
>= 0.995 and max-rel <= 8 % per tensor; label handling / dead-parameter set exact.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, grad_sample
from oracle import stonkgs_oracle as orc, weights

NAME = "cls_L2_B3_N997_K5"


def _load():
    fix = np.load(os.path.join(GOLDEN, NAME + ".npz"))
    L, B, n_kg, seed_w, seed_b, K = [int(v) for v in fix["meta"]]
    batch = {k: torch.from_numpy(fix[k]) for k in ("input_ids", "attention_mask", "token_type_ids")}
    labels = torch.from_numpy(fix["labels"])
    sd = weights.make_state_dict(n_kg, L, seed_w)
    g = torch.Generator().manual_seed(1000 + seed_w)   # oracle/make_golden.py: classifier_state
    sd["classifier.weight"] = torch.randn(K, 768, generator=g) * 0.05
    sd["classifier.bias"] = torch.randn(K, generator=g) * 0.1
    return fix, dict(layers=L, batch=B, n_kg=n_kg, K=K), batch, labels, sd, weights.make_kg_table(n_kg, seed_w)


def test_oracle_classifier_matches_reference_golden():
    fix, meta, batch, labels, sd, rows = _load()
    out, grads = orc.forward_backward_classifier(sd, orc.build_kg_table(sd, rows), dict(batch, labels=labels))
    np.testing.assert_allclose(out["logits"].detach().numpy(), fix["logits"], atol=2e-6)
    np.testing.assert_allclose(out["loss"].item(), float(fix["loss"]), atol=2e-6)
    names = [str(s) for s in fix["grad_names"]]
    assert sorted(names) == sorted(grads)
    for i, k in enumerate(names):
        g = grads[k]
        if "attention.self.key.bias" in k:
            assert g.abs().max().item() < 1e-6
            continue
        np.testing.assert_allclose(float(g.norm()), fix["grad_norms"][i], rtol=1e-4)
        s = grad_sample(g).numpy()
        np.testing.assert_allclose(s, fix["grad_samples"][i][: len(s)], atol=2e-6 * max(1.0, float(g.abs().max())) + 1e-7)


def _build(meta, sd, rows, device):
    from transformers import BertConfig
    from stonkgs_b200.finetuning import STonKGsForSequenceClassification
    cfg = BertConfig(vocab_size=28996, num_hidden_layers=meta["layers"], num_labels=meta["K"])
    model = STonKGsForSequenceClassification(None, nlp_model_type=cfg, kg_embedding_dict_path=rows, num_labels=meta["K"])
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("cls.") for k in missing), (missing, unexpected)
    return model.eval().to(device)


def test_classifier_module_tree_matches_reference_layout():
    """CPU: parameter names the fine-tuned checkpoints carry (reference stonkgs_finetuning.py:245-255)."""
    fix, meta, batch, labels, sd, rows = _load()
    model = _build(meta, sd, rows, None)
    keys = set(model.state_dict())
    assert {"classifier.weight", "classifier.bias", "bert.pooler.dense.weight", "lm_backbone.embeddings.word_embeddings.weight",
            "cls.predictions.entity_decoder.weight"} <= keys
    assert model.num_labels == meta["K"] and model.classifier.weight.shape == (meta["K"], 768)
    from stonkgs_b200._lib import StkError
    with pytest.raises(StkError):            # no CPU path
        model(**batch)


@pytest.mark.gpu
def test_classifier_forward_backward_on_gpu():
    fix, meta, batch, labels, sd, rows = _load()
    model = _build(meta, sd, rows, "cuda")
    with torch.no_grad():
        out = model(**batch, labels=labels, return_dict=True)
    np.testing.assert_allclose(out.logits.cpu().numpy(), fix["logits"], atol=4e-2)
    np.testing.assert_allclose(out.loss.item(), float(fix["loss"]), rtol=5e-3)
    tup = model(**batch)                                   # no labels, tuple form: (logits,)
    assert len(tup) == 1 and tup[0].shape == (meta["batch"], meta["K"])
    proba = model.predict_proba(batch["input_ids"].cuda(), batch["attention_mask"].cuda(), batch["token_type_ids"].cuda())
    np.testing.assert_allclose(proba.sum(1).cpu().numpy(), 1.0, atol=1e-5)
    np.testing.assert_allclose(proba.cpu().numpy(), torch.softmax(torch.from_numpy(fix["logits"]), 1).numpy(), atol=2e-2)

    model.zero_grad(set_to_none=True)
    loss = model(**batch, labels=labels)[0]
    loss.backward()
    torch.cuda.synchronize()
    ref, grads = orc.forward_backward_classifier(sd, orc.build_kg_table(sd, rows), dict(batch, labels=labels))
    np.testing.assert_allclose(loss.item(), ref["loss"].item(), rtol=5e-3)
    named = dict(model.named_parameters())
    for k, g in grads.items():
        got = named[k].grad.detach().cpu().float()
        if "attention.self.key.bias" in k:
            assert got.abs().max().item() == 0.0
            continue
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), g.reshape(1, -1)).item()
        rel = (got - g).abs().max().item() / (g.abs().max().item() + 1e-12)
        assert cos > 0.995 and rel < 0.08, (k, cos, rel)
    live = {k for k, p in named.items() if p.grad is not None}
    assert live == set(grads)                              # pre-training heads and word embeddings: no gradient
    with pytest.raises(IndexError):
        model(**batch, labels=torch.full_like(labels, meta["K"]))


@pytest.mark.gpu
def test_batched_inference_matches_rowwise():
    from stonkgs_b200.finetuning import infer_arrays, infer_iter
    fix, meta, batch, labels, sd, rows = _load()
    model = _build(meta, sd, rows, "cuda")
    ids, mask, types = (batch[k].numpy() for k in ("input_ids", "attention_mask", "token_type_ids"))
    all_at_once = infer_arrays(model, ids, mask, types, batch_size=2)
    rowwise = np.concatenate([infer_arrays(model, ids[i:i + 1], mask[i:i + 1], types[i:i + 1]) for i in range(len(ids))])
    assert np.array_equal(all_at_once, rowwise)            # batch composition does not change a row's result
    # the classifier reads the pooled [CLS] row alone: the last layer over those rows only gives the same bits
    assert np.array_equal(infer_arrays(model, ids, mask, types, batch_size=2, cls_rows_only=True), all_at_once)
    rows_iter = [dict(input_ids=ids[i].tolist(), attention_mask=mask[i].tolist(), token_type_ids=types[i].tolist())
                 for i in range(len(ids))]
    got = [p for _, p in infer_iter(model, rows_iter, batch_size=2)]
    np.testing.assert_allclose(np.asarray(got), all_at_once, atol=0)
