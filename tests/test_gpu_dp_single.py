"""GPU, one process: the data-parallel step with a world of ONE (NCCL communicator over a single B200) — the same code
path the multi-GPU runs take (bucket walk on a side stream, bf16 wire buffer, FusedAdamW reading the wire buffer with the
1 / world scale, SM reserve) checked against the plain single-GPU step.  The 2-GPU numerics (average of two half batches
= full batch, ranks bit-identical after the step) are tools/dp_check.py, run under torchrun."""
import os
import socket

import pytest
import torch

from _util import build_model, load_fixture, seeded_weights

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_world_of_one_matches_plain_step():
    import torch.distributed as dist
    from stonkgs_b200 import _lib
    from stonkgs_b200.dp import DataParallel
    from stonkgs_b200.optim import FusedAdamW
    fix, meta, batch = load_fixture("L2_B3_N3001_fullmask")
    sd, rows = seeded_weights(meta)
    dev_batch = {k: v.cuda() for k, v in batch.items()}

    def one_step(dp: bool):
        model = build_model(meta, sd, rows, "cuda")
        model.label_capacity = 64            # this fixture labels more than the default 38 positions per half
        d = DataParallel(model, bucket_mb=8.0) if dp else None
        opt = FusedAdamW(model, lr=1e-3, weight_decay=0.01, max_grad_norm=1.0)
        opt.zero_grad()
        loss = model(**dev_batch)[0]
        loss.backward()
        if dp:
            assert d.defer_unpack and d.wire_valid and len(d.buckets) >= 3
            assert _lib.load().stk_set_sm_reserve(0, 0) == 0            # finish() handed the reserved SMs back
        opt.step()
        torch.cuda.synchronize()
        flat = torch.cat([p.data.reshape(-1) for p, _ in model.grad_buffer().param_views])
        return float(loss), float(opt.grad_norm()), flat, model, d

    loss0, norm0, p0, _, _ = one_step(False)
    os.environ["NCCL_MAX_CTAS"] = "4"
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{_free_port()}", rank=0, world_size=1,
                            device_id=torch.device("cuda", 0))
    try:
        loss1, norm1, p1, model, d = one_step(True)
        assert d.sm_reserve == 4
        assert loss1 == loss0
        assert abs(norm1 - norm0) < 2e-3 * norm0                        # norm of the bf16-rounded gradient
        diff = (p1 - p0).abs()
        # AdamW's first step moves every weight by ~lr * sign(g): bf16 rounding of g only matters where g is ~0
        assert diff.max().item() <= 2.1e-3 and (diff > 1e-5).float().mean().item() < 2e-3
        # the averaged gradient written back on request == the local gradient through one bf16 rounding
        gb = model.grad_buffer()
        local = gb.flat.clone()
        d.wire_valid = True
        d.materialize_grads()
        torch.cuda.synchronize()
        assert torch.equal(gb.flat, local.bfloat16().float())
    finally:
        dist.destroy_process_group()
        os.environ.pop("NCCL_MAX_CTAS", None)
