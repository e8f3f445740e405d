"""Training-mode dropout.

torch's dropout masks cannot be reproduced outside torch, so parity is defined with INJECTED masks:
  (CPU)  the oracle with a DropSpec == the reference's own module in train() with ``torch.nn.functional.dropout``
         replaced by the same DropSpec, site ids assigned in the reference's call order — pins WHERE the oracle (and
         therefore the CUDA path) applies dropout;
  (GPU)  the CUDA model in train() == the oracle with the DropSpec of the seed the model reports.
"""
import numpy as np
import pytest
import torch

from _util import build_model, load_fixture, seeded_weights
from oracle import dropout_oracle as do
from oracle import ref_shim, stonkgs_oracle as orc


def test_keep_mask_statistics_and_determinism():
    m = do.keep_mask(123, 7, 4096, 768, 13)
    assert abs((1 - m.mean()) - 13 / 128) < 1.5e-3
    assert np.array_equal(m, do.keep_mask(123, 7, 4096, 768, 13))
    assert (m != do.keep_mask(124, 7, 4096, 768, 13)).mean() > 0.1
    assert (m != do.keep_mask(123, 8, 4096, 768, 13)).mean() > 0.1
    assert abs(m.mean(0).std()) < 0.01 and abs(m.mean(1).std()) < 0.02      # no row / column structure
    assert do.keep_mask(1, 1, 8, 16, 0).all()
    # the eight decisions that share one hash (two words, four 7-bit fields each) are pairwise uncorrelated, and so are
    # vertical neighbours (consecutive row keys)
    d = (~m).astype(np.float64)
    blk = d.reshape(4096, 96, 8)
    p = d.mean()
    for a in range(8):
        for b in range(a + 1, 8):
            joint = (blk[:, :, a] * blk[:, :, b]).mean()
            assert abs(joint - p * p) < 1.2e-3, (a, b, joint, p * p)
    assert abs((d[1:] * d[:-1]).mean() - p * p) < 1e-3
    # known answers of the decision words (pin the bit layout shared with csrc/stk_rng.cuh)
    w0, w1 = do.drop_words(np.uint32(0x12345678), np.uint32(5))
    assert (int(w0), int(w1)) == (int(do.lowbias32(np.uint32((0x12345678 + 5 * 0x9E3779B9) & 0xFFFFFFFF))),
                                  int(((int(w0) * 0x9E3779B1) >> 32) ^ ((int(w0) * 0x9E3779B1) & 0xFFFFFFFF)))


@pytest.mark.skipif(not ref_shim.reference_available(), reason="reference tree only exists in the dev container")
def test_oracle_dropout_sites_match_reference_in_train_mode():
    fix, meta, batch = load_fixture("L2_B2_N997")
    sd, rows = seeded_weights(meta)
    L = meta["layers"]
    spec = do.DropSpec(seed=4242, p_hidden=0.1, p_attn=0.1)
    # the reference's dropout calls, in execution order: LM backbone (encoder 0) then joint encoder (1), each
    # embeddings (HF:110), then per layer attention probabilities (:132), attention output (:297), FFN output (:355)
    order = []
    for enc in (0, 1):
        order.append((do.site_embeddings(enc), False))
        for li in range(L):
            order += [(do.site_attention(enc, li), True), (do.site_attn_out(enc, li), False), (do.site_ffn_out(enc, li), False)]
    calls = {"i": 0}
    real = torch.nn.functional.dropout

    def injected(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        site, is_attn = order[calls["i"]]
        calls["i"] += 1
        assert is_attn == (input.dim() == 4), (calls["i"], input.shape)
        assert abs(p - 0.1) < 1e-9
        return spec(site, input, attention=is_attn)

    ref = ref_shim.load_reference(sd, rows, L).train()
    torch.nn.functional.dropout = injected
    try:
        out = ref(**batch, return_dict=True)
        out.loss.backward()
    finally:
        torch.nn.functional.dropout = real
        ref.eval()
    assert calls["i"] == len(order)
    o, grads = orc.forward_backward(sd, orc.build_kg_table(sd, rows), batch, drop=spec)
    assert torch.equal(o["pooler_output"], out.pooler_output.detach())
    assert abs(o["loss"].item() - out.loss.item()) < 5e-6
    ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    for k, g in grads.items():
        if "attention.self.key.bias" in k:
            continue
        assert (g - ref_grads[k]).abs().max().item() <= 2e-6 * (ref_grads[k].abs().max().item() + 1e-30) + 1e-9, k
    eval_out = orc.forward(sd, orc.build_kg_table(sd, rows), **batch)
    assert (eval_out["pooler_output"] - o["pooler_output"]).abs().max() > 1e-3     # dropout really changed the pass


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["L2_B2_N997", "L2_B3_N3001_fullmask"])
def test_train_mode_matches_oracle_with_same_masks(name):
    fix, meta, batch = load_fixture(name)
    sd, rows = seeded_weights(meta)
    model = build_model(meta, sd, rows, "cuda").train()
    model.stk_dropout_seed = 777
    model.zero_grad(set_to_none=True)
    out = model(**batch, return_dict=True)
    seed = model.last_dropout_seed
    assert seed is not None
    out.loss.backward()
    torch.cuda.synchronize()
    spec = do.DropSpec(seed, model.config.hidden_dropout_prob, model.config.attention_probs_dropout_prob)
    ref, grads = orc.forward_backward(sd, orc.build_kg_table(sd, rows), batch, drop=spec)
    np.testing.assert_allclose(out.pooler_output.cpu().numpy(), ref["pooler_output"].detach().numpy(), atol=1e-1)
    assert (out.pooler_output.cpu() - ref["pooler_output"].detach()).abs().mean() < 2e-2
    np.testing.assert_allclose(out.loss.item(), ref["loss"].item(), rtol=4e-3)
    named = dict(model.named_parameters())
    for k, g in grads.items():
        got = named[k].grad.detach().cpu().float()
        if "attention.self.key.bias" in k:
            assert got.abs().max().item() == 0.0
            continue
        cos = torch.nn.functional.cosine_similarity(got.reshape(1, -1), g.reshape(1, -1)).item()
        rel = (got - g).abs().max().item() / (g.abs().max().item() + 1e-12)
        assert cos > 0.99 and rel < 0.12, (k, cos, rel)
    # the eval() forward is untouched by all of this, and a second train() step draws new masks
    model.eval()
    with torch.no_grad():
        ev = model(**batch, return_dict=True)
    np.testing.assert_allclose(ev.pooler_output.cpu().numpy(), fix["pooler_output"], atol=8e-2)
    model.train()
    out2 = model(**batch, return_dict=True)
    assert model.last_dropout_seed != seed and abs(out2.loss.item() - out.loss.item()) > 1e-4
    model.stk_dropout = False
    out3 = model(**batch, return_dict=True)
    np.testing.assert_allclose(out3.loss.item(), float(fix["loss"]), rtol=2e-3)   # dropout off == eval numerics
