"""GPU: every kernel of libstk.so against plain PyTorch fp32 references of the same op, through the
C ABI (tools/gpu_probe.py holds the case bodies; each returns per-check error reports)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu


def _run(case):
    import gpu_probe
    reports = gpu_probe.CASES[case]()
    bad = [r for r in reports if not r.get("ok", False)]
    assert not bad, bad
    return reports


@pytest.mark.parametrize("case", ["elementwise", "gemm_basic", "gemm_epilogues", "gemm_majors", "gemm_ln", "gemm_ce", "cls_head", "dropout", "attn",
                                  "attn_bwd"])
def test_kernel_group(case):
    _run(case)


@pytest.mark.parametrize("case", ["gemm_basic", "gemm_epilogues", "gemm_majors", "gemm_ce"])
def test_gemm_with_dynamic_tile_scheduling(case):
    """The same GEMM cases with the cluster-launch-control scheduler (stk_set_gemm_dynamic): one CTA pair per tile is
    launched, running pairs cancel pending ones and take over their tiles; results must not depend on the schedule."""
    from stonkgs_b200 import _lib
    lib = _lib.load()
    prev = lib.stk_set_gemm_dynamic(1)
    try:
        _run(case)
    finally:
        lib.stk_set_gemm_dynamic(prev)


def test_product_never_touches_oracle():
    """The product path must not import anything from oracle/ (only tests / bench may)."""
    import subprocess
    code = ("import sys, torch; from stonkgs_b200 import model, training, embeddings, engine, ops; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def test_dropout_inside_the_layernorm_epilogue():
    """STK_EPI_BIAS_DROP_RESID_LN (train(): z = drop(x W^T + b) + r, y = LN(z)) against the unfused pair it replaces
    (bias GEMM, then the dropout + residual + LayerNorm row kernel): the keep decisions are the same pure function of
    (seed, site, row, column), so the dropped positions — where z is exactly the residual — must coincide bit for bit;
    values agree to bf16 rounding (the fused form does not round the dense output to bf16 before the sum)."""
    from stonkgs_b200 import ops
    torch.manual_seed(0)
    for M, K in ((1024, 768), (640, 3072)):
        x = torch.randn(M, K, device="cuda").bfloat16()
        w = (torch.randn(768, K, device="cuda") * K ** -0.5).bfloat16()
        b = torch.randn(768, device="cuda") * 0.1
        r = torch.randn(M, 768, device="cuda").bfloat16()
        g = 1 + 0.1 * torch.randn(768, device="cuda")
        beta = 0.1 * torch.randn(768, device="cuda")
        d = ops.Drop(seed=1234567, site=69, p=0.1)
        dense = ops.linear(x, w, b)
        y0, z0, m0, s0 = ops.dropout_resid_ln(dense, r, g, beta, d, save_for_backward=True)
        y1, z1, m1, s1 = ops.linear_resid_ln(x, w, b, r, g, beta, save_for_backward=True, drop=d)
        dropped0, dropped1 = z0 == r, z1 == r
        frac = dropped1.float().mean().item()
        assert abs(frac - 13 / 128) < 5e-3, frac
        # (a tiny dense value can also leave z == r after rounding: compare the decisions where it cannot)
        big = dense.float().abs() > 0.25
        assert torch.equal(dropped0[big], dropped1[big])
        assert (dropped0 != dropped1).float().mean().item() < 2e-3
        torch.testing.assert_close(z1.float(), z0.float(), atol=3e-2, rtol=2e-2)
        torch.testing.assert_close(y1.float(), y0.float(), atol=4e-2, rtol=2e-2)
        torch.testing.assert_close(m1, m0, atol=2e-3, rtol=0)
        torch.testing.assert_close(s1, s0, atol=0, rtol=5e-3)
        # fp32 reference of the fused form
        keep = (~dropped1).float() * (128.0 / 115.0)
        zf = (x.float() @ w.float().T + b) * keep + r.float()
        torch.testing.assert_close(z1.float(), zf, atol=2e-2, rtol=1e-2)
        yf = torch.nn.functional.layer_norm(z1.float(), (768,), g, beta, 1e-12)
        torch.testing.assert_close(y1.float(), yf, atol=3e-2, rtol=1e-2)
