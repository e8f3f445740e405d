"""GPU: fused clip + AdamW against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW (HF Trainer's
defaults for the reference's pre-training, stonkgs_pretraining.py:171-193) over several real steps."""
import pytest
import torch

from _util import build_model, load_fixture, seeded_weights

pytestmark = pytest.mark.gpu


def test_fused_adamw_matches_torch():
    from stonkgs_b200.optim import FusedAdamW
    fix, meta, batch = load_fixture("L2_B2_N997")
    sd, rows = seeded_weights(meta)
    model = build_model(meta, sd, rows, "cuda")
    opt = FusedAdamW(model, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
    live = [(n, p) for n, p in model.named_parameters() if any(p is q for q, _ in model.grad_buffer().param_views)]
    assert len(live) == 46
    shadow = {n: p.detach().clone().requires_grad_(True) for n, p in live}
    ref_opt = torch.optim.AdamW(list(shadow.values()), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01)
    for step in range(3):
        opt.zero_grad()
        loss = model(**batch)[0]
        loss.backward()
        for n, p in live:                       # same gradients on the torch side
            shadow[n].grad = p.grad.detach().clone()
        norm_ref = torch.nn.utils.clip_grad_norm_(list(shadow.values()), 1.0)
        ref_opt.step()
        opt.step()
        torch.testing.assert_close(opt.grad_norm().reshape(()), norm_ref.reshape(()), rtol=1e-4, atol=1e-6)
        for n, p in live:
            torch.testing.assert_close(p.detach(), shadow[n].detach(), rtol=2e-5, atol=2e-7, msg=lambda m: f"{n} step {step}: {m}")
    # the bf16 GEMM copies were refreshed by the optimizer pass itself (no recast needed)
    st = model._dev_state
    l0 = model.bert.encoder.layer[0]
    assert torch.equal(st["bert"].layers[0].w1, l0.intermediate.dense.weight.detach().bfloat16())
    assert torch.equal(st["bert"].layers[0].wqkv[768:1536], l0.attention.self.key.weight.detach().bfloat16())
    assert torch.equal(st["bert"].layers[0].bqkv[:768], l0.attention.self.query.bias.detach())
    assert torch.equal(st["heads"].w_ent, model.cls.predictions.entity_decoder.weight.detach().bfloat16())
    # and the loss goes down when the same batch is fitted
    l0_ = model(**batch)[0].item()
    for _ in range(5):
        opt.zero_grad()
        model(**batch)[0].backward()
        opt.step()
    assert model(**batch)[0].item() < l0_
    # state round trip
    sd2 = opt.state_dict()
    opt2 = FusedAdamW(model, lr=1e-3)
    opt2.load_state_dict(sd2)
    assert opt2._step == opt._step and torch.equal(opt2.exp_avg, opt.exp_avg)
