"""TransE variant: ``TransESTonKGsForPreTraining`` (transestonkgs_model.py) = the STonKGs model on a
256 + 4 token sequence with 260 positions.

CPU: the oracle (``forward(..., text_len=256)`` on a [B, 260] batch) against the golden vectors produced by the
reference's OWN TransE class (tests/golden/transe_*.npz, oracle/make_golden.py), and live against that class when
/root/reference is present; the drop-in's module tree against the reference's state-dict keys.
GPU: the drop-in (384 padded rows per pair inside, 260 positions outside) against fixture and oracle, same stated
tolerances as tests/test_gpu_e2e.py.
"""
import os

import numpy as np
import pytest
import torch

from _util import GOLDEN, BATCH_KEYS, grad_sample
from oracle import ref_shim, stonkgs_oracle as orc, weights

NAME = "transe_L2_B3_N499"


def _load():
    fix = np.load(os.path.join(GOLDEN, NAME + ".npz"))
    L, B, n_kg, seed_w, seed_b, _ = [int(v) for v in fix["meta"]]
    batch = {k: torch.from_numpy(fix[k]) for k in BATCH_KEYS}
    sd = weights.make_state_dict(n_kg, L, seed_w, joint_max_pos=260)
    rows = weights.make_kg_table(n_kg, seed_w)
    return fix, dict(layers=L, batch=B, n_kg=n_kg), batch, sd, rows


def _build(meta, sd, rows, device=None):
    from transformers import BertConfig
    from stonkgs_b200.model import TransESTonKGsForPreTraining
    model = TransESTonKGsForPreTraining(None, BertConfig(vocab_size=28996, num_hidden_layers=meta["layers"]), rows)
    model.load_state_dict(sd, strict=True)
    model.eval()
    return model.to(device) if device is not None else model


def test_oracle_matches_reference_transe_golden():
    fix, meta, batch, sd, rows = _load()
    assert batch["input_ids"].shape == (meta["batch"], 260) and batch["ent_masked_lm_labels"].shape == (meta["batch"], 4)
    out, grads = orc.forward_backward(sd, orc.build_kg_table(sd, rows), batch, text_len=256)
    r = [int(v) for v in fix["rows"]]
    assert out["sequence_output"].shape == (meta["batch"], 260, 768)
    np.testing.assert_allclose(out["pooler_output"].detach().numpy(), fix["pooler_output"], atol=2e-6, rtol=0)
    np.testing.assert_allclose(out["sequence_output"].detach()[:, r].numpy(), fix["sequence_output_rows"], atol=1e-5, rtol=0)
    np.testing.assert_allclose(out["loss"].item(), float(fix["loss"]), rtol=2e-6)
    np.testing.assert_allclose(out["elm_loss"].item(), float(fix["elm_loss"]), rtol=2e-6)
    np.testing.assert_allclose(out["entity_lse"].detach().numpy(), fix["entity_lse"], atol=1e-5)
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
    assert grads["bert.embeddings.position_embeddings.weight"].shape == (260, 768)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the dev container")
def test_oracle_bitwise_vs_live_transe_reference_and_key_layout():
    from stonkgs_b200 import synthetic
    n_kg, layers = 211, 1
    sd = weights.make_state_dict(n_kg, layers, 12, joint_max_pos=260)
    rows = weights.make_kg_table(n_kg, 12)
    batch = synthetic.make_batch(2, n_kg, seed=4, kg_len=4)
    ref = ref_shim.load_reference_transe(sd, rows, layers)
    with torch.no_grad():
        r = ref(**batch, return_dict=True)
        o = orc.forward(sd, orc.build_kg_table(sd, rows), **batch, text_len=256)
    assert torch.equal(r.pooler_output, o["pooler_output"]) and torch.equal(r.hidden_states, o["sequence_output"])
    assert abs(r.loss.item() - o["loss"].item()) < 5e-6
    # a pair batch without any entity label: CrossEntropyLoss over zero rows is NaN on both sides (torch semantics)
    nolab = dict(batch, ent_masked_lm_labels=torch.full((2, 4), -100))
    with torch.no_grad():
        assert torch.isnan(ref(**nolab, return_dict=True).loss)
        assert torch.isnan(orc.forward(sd, orc.build_kg_table(sd, rows), **nolab, text_len=256)["loss"])
    # the drop-in's module tree carries exactly the reference's checkpoint keys and shapes
    mine = _build(dict(layers=layers), sd, rows)
    ref_sd, my_sd = ref.state_dict(), mine.state_dict()
    assert sorted(ref_sd) == sorted(my_sd)
    assert all(ref_sd[k].shape == my_sd[k].shape for k in ref_sd)
    assert mine.cls.predictions.text_part_length == ref.cls.predictions.text_part_length == 256
    assert mine.config.max_position_embeddings == ref.config.max_position_embeddings == 260
    assert mine.config.kg_vocab_size == ref.config.kg_vocab_size == n_kg


def test_shape_contract_errors():
    from stonkgs_b200 import ops
    from stonkgs_b200._lib import StkError
    assert ops.SeqShape(256, 260).seq_pad == 384 and ops.SeqShape(256, 260).kg_len == 4
    assert ops.STONKGS_SHAPE.seq_pad == 512
    with pytest.raises(StkError):
        ops.SeqShape(256, 600)
    with pytest.raises(StkError):
        ops.SeqShape(256, 256)


@pytest.mark.gpu
def test_transe_forward_backward_on_gpu():
    fix, meta, batch, sd, rows = _load()
    model = _build(meta, sd, rows, "cuda")
    B = meta["batch"]
    with torch.no_grad():
        out = model(**batch, return_dict=True)
    r = [int(v) for v in fix["rows"]]
    assert out.hidden_states.shape == (B, 260, 768) and out.pooler_output.shape == (B, 768)
    np.testing.assert_allclose(out.pooler_output.cpu().numpy(), fix["pooler_output"], atol=8e-2)
    assert np.abs(out.pooler_output.cpu().numpy() - fix["pooler_output"]).mean() < 1.5e-2
    np.testing.assert_allclose(out.hidden_states[:, r].cpu().numpy(), fix["sequence_output_rows"], atol=1e-1)
    np.testing.assert_allclose(out.loss.item(), float(fix["loss"]), rtol=2e-3)
    mlm, elm, nsp = [float(v) for v in model._last_loss_parts]
    np.testing.assert_allclose(mlm, float(fix["mlm_loss"]), rtol=2e-3)
    np.testing.assert_allclose(elm, float(fix["elm_loss"]), rtol=2e-3)
    np.testing.assert_allclose(out.seq_relationship_logits.cpu().numpy(), fix["seq_relationship_logits"], atol=3e-2)

    # bit-exact gather of the four KG rows and of the LM rows into the padded layout; padding rows are zero
    from stonkgs_b200 import engine, ops
    st = model._device_state(False)
    ids = batch["input_ids"].cuda()
    lm_hidden = engine.lm_backbone_fwd(st["lm"], ids[:, :256])
    x, _, _, emb = ops.embed_joint_ln(ids, None, lm_hidden, model.kg_table, st["bert"].pos, st["bert"].type_emb,
                                      st["bert"].emb_g, st["bert"].emb_b, want_inputs_embeds=True, shape=model.seq_shape)
    emb = emb.view(B, 384, 768)
    assert torch.equal(emb[:, 256:260], model.kg_table[ids[:, 256:]])
    assert torch.equal(emb[:, :256], lm_hidden.view(B, 256, 768).float())
    assert emb[:, 260:].abs().max().item() == 0.0 and x.view(B, 384, 768)[:, 260:].float().abs().max().item() == 0.0

    # dense logits on request keep the reference's shapes ([B,256,V] and [B,4,N])
    model.return_prediction_logits = True
    with torch.no_grad():
        text, ent = model(**batch, return_dict=True).prediction_logits
    model.return_prediction_logits = False
    assert text.shape == (B, 256, 28996) and ent.shape == (B, 4, meta["n_kg"])
    sel = (batch["ent_masked_lm_labels"].reshape(-1) != -100).cuda()
    np.testing.assert_allclose(torch.logsumexp(ent.reshape(-1, meta["n_kg"])[sel], -1).cpu().numpy(), fix["entity_lse"], atol=2e-2)

    # backward against the oracle
    model.zero_grad(set_to_none=True)
    loss = model(**batch)[0]
    loss.backward()
    torch.cuda.synchronize()
    ref, grads = orc.forward_backward(sd, orc.build_kg_table(sd, rows), batch, text_len=256)
    np.testing.assert_allclose(loss.item(), ref["loss"].item(), rtol=2e-3)
    named = dict(model.named_parameters())
    for k, g in grads.items():
        got = named[k].grad.detach().cpu().float()
        assert got.shape == g.shape, k
        if "attention.self.key.bias" in k:
            assert got.abs().max().item() == 0.0
            continue
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), g.reshape(1, -1)).item()
        rel = (got - g).abs().max().item() / (g.abs().max().item() + 1e-12)
        assert cos > 0.995 and rel < 0.08, (k, cos, rel)
    # a batch without entity labels gives NaN like the reference (CrossEntropyLoss over zero rows)
    nolab = dict(batch, ent_masked_lm_labels=torch.full((B, 4), -100))
    with torch.no_grad():
        assert torch.isnan(model(**nolab)[0]).item()
    # train() mode (dropout on the padded 384-row layout): finite loss, gradients flow, eval numerics with dropout off
    model.train()
    model.zero_grad(set_to_none=True)
    l1 = model(**batch)[0]
    l1.backward()
    torch.cuda.synchronize()
    assert torch.isfinite(l1).item() and abs(l1.item() - loss.item()) > 1e-4
    assert all(torch.isfinite(p.grad).all().item() for p in model.parameters() if p.grad is not None)
    model.stk_dropout = False
    np.testing.assert_allclose(model(**batch)[0].item(), float(fix["loss"]), rtol=2e-3)
    model.eval()
    # wrong width is refused loudly
    from stonkgs_b200._lib import StkError
    with pytest.raises(StkError):
        model(input_ids=torch.zeros((1, 512), dtype=torch.long))
