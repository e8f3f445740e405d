"""GPU: the device-side pieces of the pre-training heads (no host syncs inside the training step) and head_mask.

* stk_compact_labels against torch.nonzero (bit-exact indices, row-major order, padding entries, flags);
* labels resident on the device (fixed-capacity path) give the same loss and gradients as labels on the host (exact
  capacity) — and as the oracle;
* out-of-range labels / too many labelled positions are flagged on the device and raised at the deferred check;
* head_mask: the reference hands it to the joint encoder (stonkgs_model.py:158,209); compared with the oracle's
  restatement of transformers 4.x (``attention_probs * head_mask``).
"""
import numpy as np
import pytest
import torch

from _util import build_model, load_fixture, seeded_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small():
    fix, meta, batch = load_fixture("L2_B3_N3001_fullmask")
    sd, rows = seeded_weights(meta)
    model = build_model(meta, sd, rows, "cuda")
    model.label_capacity = 64     # this fixture labels more positions per half than the reference's 38
    return meta, batch, sd, rows, model


def test_compact_labels_matches_nonzero():
    from stonkgs_b200 import ops
    g = torch.Generator().manual_seed(3)
    for B, width, pitch, off, density in ((7, 256, 512, 256, 0.15), (64, 256, 512, 0, 0.15), (5, 4, 384, 256, 0.5), (3, 256, 512, 0, 0.0)):
        labels = torch.full((B, width), -100, dtype=torch.int64)
        pick = torch.rand((B, width), generator=g) < density
        labels[pick] = torch.randint(0, 1000, (int(pick.sum()),), generator=g)
        n = int(pick.sum())
        cap = n + 13
        err = torch.zeros(1, dtype=torch.int32, device="cuda")
        rows, labs, count = ops.compact_labels(labels.cuda(), pitch, off, 1000, cap, err)
        pos = torch.nonzero(pick)
        want_rows = (pos[:, 0] * pitch + pos[:, 1] + off).to(torch.int32)
        assert int(count) == n and int(err) == 0
        assert torch.equal(rows[:n].cpu(), want_rows) and torch.equal(labs[:n].cpu(), labels[pick].to(torch.int32))
        assert (rows[n:] == -1).all() and (labs[n:] == -1).all()
    # flags: label outside the vocabulary (bit 1), more labelled positions than capacity (bit 2)
    labels = torch.full((2, 256), -100, dtype=torch.int64)
    labels[0, :10] = torch.arange(10)
    labels[1, 5] = 1000
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    rows, labs, count = ops.compact_labels(labels.cuda(), 512, 0, 1000, 8, err)
    assert int(err) == 2 | 4 and int(count) == 10 and torch.equal(labs.cpu(), torch.arange(8, dtype=torch.int32))


def test_device_labels_equal_host_labels(small):
    meta, batch, sd, rows, model = small
    named = dict(model.named_parameters())

    def run(b):
        model.zero_grad(set_to_none=True)
        loss = model(**b)[0]
        loss.backward()
        torch.cuda.synchronize()
        model._raise_on_bad_ids()
        return loss.item(), {k: p.grad.clone() for k, p in named.items() if p.grad is not None}

    loss_h, g_h = run(batch)                                      # labels on the host: exact capacity
    loss_d, g_d = run({k: v.cuda() for k, v in batch.items()})    # labels on the device: padded, fixed capacity (64)
    assert loss_h == loss_d                                       # the forward is deterministic: bit-identical
    # padding rows contribute exactly nothing; what differs is fp32 summation order in the heads (split-K factors follow
    # the row count, reduce-adds land in arrival order): an fp32 ulp there can flip a bf16 rounding of the gradient that
    # travels down the trunk, i.e. the run-to-run noise of two identical steps (measured: <= 1e-3 of a tensor's max)
    for k in g_h:
        scale = g_h[k].abs().max().item() + 1e-20
        assert (g_h[k] - g_d[k]).abs().max().item() <= 5e-3 * scale, k
    model.zero_grad(set_to_none=True)


def test_label_flags_raise_at_the_deferred_check(small):
    from stonkgs_b200 import StkError
    meta, batch, sd, rows, model = small
    dev_batch = {k: v.cuda() for k, v in batch.items()}
    bad = {k: v.clone() for k, v in dev_batch.items()}
    bad["ent_masked_lm_labels"][0, 3] = meta["n_kg"]              # entity labels live in [0, kg_vocab_size)
    model(**bad)[0].backward()
    with pytest.raises(IndexError):
        model._raise_on_bad_ids()
    with pytest.raises(IndexError):                               # host labels: checked before anything is launched
        model(**{k: v.cpu() for k, v in bad.items()})
    model.label_capacity = 8                                      # fewer rows than the batch labels
    try:
        model(**dev_batch)[0].backward()
        with pytest.raises(StkError):
            model._raise_on_bad_ids()
    finally:
        model.label_capacity = 64
    bad = {k: v.clone() for k, v in dev_batch.items()}
    bad["next_sentence_labels"][1] = 2
    with torch.no_grad(), pytest.raises(IndexError):
        model(**bad)
    model.zero_grad(set_to_none=True)
    with torch.no_grad():
        assert torch.isfinite(model(**dev_batch)[0])              # flags do not stick


def test_head_mask_matches_oracle(small):
    from oracle import stonkgs_oracle as orc
    meta, batch, sd, rows, model = small
    table = orc.build_kg_table(sd, rows)
    hm = torch.ones(meta["layers"], 12)
    hm[0, 3] = 0.0
    hm[1, 7] = 0.5
    hm[1, 0] = 0.0
    with torch.no_grad():
        base = model(**batch, return_dict=True)
        ones = model(**batch, return_dict=True, head_mask=torch.ones(12))
    assert torch.equal(base.pooler_output, ones.pooler_output)     # an all-ones mask is the identity, bit for bit
    model.zero_grad(set_to_none=True)
    out = model(**batch, return_dict=True, head_mask=hm)
    out.loss.backward()
    torch.cuda.synchronize()
    ref, grads = orc.forward_backward(sd, table, batch, head_mask=hm)
    assert (out.pooler_output - base.pooler_output).abs().max().item() > 1e-3   # the mask does something
    np.testing.assert_allclose(out.pooler_output.detach().cpu().numpy(), ref["pooler_output"].detach().numpy(), atol=5e-2)
    np.testing.assert_allclose(out.loss.item(), ref["loss"].item(), rtol=2e-3)
    named = dict(model.named_parameters())
    for k in ("bert.encoder.layer.0.attention.self.value.weight", "bert.encoder.layer.1.attention.output.dense.weight",
              "bert.encoder.layer.0.attention.self.query.weight", "bert.embeddings.position_embeddings.weight"):
        got, g = named[k].grad.detach().cpu(), grads[k]
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), g.reshape(1, -1)).item()
        assert cos > 0.999, (k, cos)
    # the value rows of a fully masked head get no gradient at all
    gv = named["bert.encoder.layer.0.attention.self.value.weight"].grad[3 * 64:4 * 64]
    assert gv.abs().max().item() == 0.0 and grads["bert.encoder.layer.0.attention.self.value.weight"][3 * 64:4 * 64].abs().max() == 0
    model.zero_grad(set_to_none=True)
