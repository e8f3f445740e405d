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
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
