"""2-GPU check of the data-parallel path (run under torchrun): averaged gradients of two half
batches == single-process gradients of the full batch (equal labelled-row counts per sample)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from stonkgs_b200 import synthetic  # noqa: E402
from stonkgs_b200.dp import DataParallel  # noqa: E402


def main():
    world, rank, local = bench.dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n_kg, layers, per = 5003, 2, 2
    model = bench.build_model(dev, layers, n_kg, seed=0)
    full = synthetic.make_batch(per * world, n_kg, seed=7)
    mine = {k: v[rank * per:(rank + 1) * per] for k, v in full.items()}
    # reference: full batch, no DP, on every rank (same weights by construction)
    model.zero_grad(set_to_none=True)
    model(**full)[0].backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    model.zero_grad(set_to_none=True)
    # ranks start from DIFFERENT weights with their device-side copies already built (the forward above built them):
    # the broadcast must reach the bf16 GEMM copies too, not only the fp32 masters
    if rank != 0:
        with torch.no_grad():
            for p in model.bert.parameters():
                p.add_(0.01 * rank)
        model(**mine)   # rebuilds the bf16 copies from the perturbed weights
    DataParallel(model, dist.group.WORLD)
    loss = model(**mine)[0]
    loss.backward()
    torch.cuda.synchronize()
    worst = 0.0
    for n, p in model.named_parameters():
        if p.grad is None:
            continue
        denom = ref[n].abs().max().item() + 1e-12
        if "key.bias" in n:
            continue
        worst = max(worst, (p.grad - ref[n]).abs().max().item() / denom)
    # every rank must hold identical gradients after the all-reduce
    flat = model.grad_buffer().flat
    other = flat.clone()
    dist.broadcast(other, src=0)
    same = bool(torch.equal(other, flat))
    print(f"rank {rank}: dp vs full-batch worst rel grad diff = {worst:.4f}; identical across ranks = {same}; "
          f"buckets = {len(model._dp.buckets)}", flush=True)
    assert worst < 3e-2 and same

    # ---- with FusedAdamW attached: no unpack pass, the optimizer consumes the summed bf16 wire buffer ----------------
    from stonkgs_b200.optim import FusedAdamW
    opt = FusedAdamW(model, lr=1e-3, weight_decay=0.0, max_grad_norm=1.0)
    assert model._dp.defer_unpack
    opt.zero_grad()
    model(**mine)[0].backward()
    assert model._dp.wire_valid
    local = model.grad_buffer().flat.clone()          # param.grad now holds the rank-LOCAL gradient
    wire = model._dp._wire.float() / world
    torch.cuda.synchronize()
    worst2 = 0.0
    for n, p in model.named_parameters():
        if p.grad is None or "key.bias" in n:
            continue
        off = (p.grad.data_ptr() - model.grad_buffer().flat.data_ptr()) // 4
        got = wire[off:off + p.numel()].view_as(p)
        worst2 = max(worst2, (got - ref[n]).abs().max().item() / (ref[n].abs().max().item() + 1e-12))
    # clip norm from the wire buffer == norm of the averaged gradient
    opt.step()
    torch.cuda.synchronize()
    norm_wire = float(opt.grad_norm())
    norm_ref = float(torch.sqrt(sum((g.float() ** 2).sum() for g in ref.values())))
    flatp = torch.cat([p.data.reshape(-1) for p, _ in model.grad_buffer().param_views])
    other = flatp.clone()
    dist.broadcast(other, src=0)
    same_p = bool(torch.equal(other, flatp))
    losses = []
    for _ in range(4):
        opt.zero_grad()
        loss = model(**mine)[0]
        loss.backward()
        opt.step()
        losses.append(float(loss))
    print(f"rank {rank}: fused-optimizer path: wire vs full-batch worst rel = {worst2:.4f}; grad norm {norm_wire:.4f} vs "
          f"{norm_ref:.4f}; params identical across ranks after step = {same_p}; losses {[round(x, 3) for x in losses]}",
          flush=True)
    assert worst2 < 3e-2 and same_p and abs(norm_wire - norm_ref) < 2e-2 * norm_ref and losses[-1] < losses[0]
    assert not torch.equal(local, wire)               # (sanity: local and averaged gradients differ)
    model._dp.materialize_grads()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
