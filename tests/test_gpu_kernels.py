"""GPU: every kernel of libstk.so against plain PyTorch fp32 references of the same op, through the
C ABI (tools/gpu_probe.py holds the case bodies; each returns per-check error reports)."""
import os
import sys

import pytest

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


def test_product_never_touches_oracle():
    """The product path must not import anything from oracle/ (only tests / bench may)."""
    import subprocess
    code = ("import sys, torch; from stonkgs_b200 import model, training, embeddings, engine, ops; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'")
    subprocess.run([sys.executable, "-c", code], check=True, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
