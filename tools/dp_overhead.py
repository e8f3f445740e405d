"""Where the data-parallel step's overhead goes, on ONE GPU (NCCL world of one): the pre-training step of bench.py with
(a) no DataParallel object, (b) the bucket walk only (STK_DP_DEBUG_SKIP_AR=1 STK_DP_DEBUG_SKIP_PACK=1), with the optimizer
reading the wire buffer or the fp32 gradients.  Prints ms per step for each variant."""
import os
import socket
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("STK_DP_DEBUG_SKIP_AR", "1")
os.environ.setdefault("STK_DP_DEBUG_SKIP_PACK", "1")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from stonkgs_b200 import synthetic  # noqa: E402
from stonkgs_b200.dp import DataParallel  # noqa: E402
from stonkgs_b200.optim import FusedAdamW  # noqa: E402


def main():
    world, rank, local = bench.dist_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(local)
    if world > 1:      # under torchrun: the same variants with several ranks side by side (still no collective)
        dist.init_process_group("nccl", device_id=dev)
    else:
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=0, world_size=1, device_id=dev)
    model = bench.build_model(dev, 12).train()
    batches = [{k: v.to(dev) for k, v in synthetic.make_batch(64, bench.N_KG, seed=200 + i).items()} for i in range(4)]

    def timed(opt, steps=20):
        def step(i):
            opt.zero_grad()
            model(**batches[i % 4])[0].backward()
            opt.step()
        for i in range(5):
            step(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            step(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    res = {}
    opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
    res["no DataParallel"] = timed(opt)
    dp = DataParallel(model)
    opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
    res["bucket walk only, optimizer reads the wire buffer"] = timed(opt)
    with dp.no_sync():
        res["DataParallel.no_sync()"] = timed(opt)
    dp.defer_unpack = False
    opt._segs_wire_dev = None
    res["bucket walk only, optimizer reads fp32 gradients (+ unpack kernels)"] = timed(opt)
    dp.overlap = False
    res["bucket walk at the end of backward, fp32 optimizer"] = timed(opt)
    dp.overlap = True
    dp.sm_reserve = 0
    res["bucket walk only, no SM reserve"] = timed(opt)
    if world > 1:
        # bench.py's harness: a barrier (an NCCL collective) on both sides of the timed region
        dp.sm_reserve = 4
        dp.defer_unpack = True
        opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)

        def timed_b(steps=10):
            def step(i):
                opt.zero_grad()
                model(**batches[i % 4])[0].backward()
                opt.step()
            step(0)
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(i)
            e1.record()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps

        res["[barriers] bucket walk only, wire optimizer"] = timed_b()
        with dp.no_sync():
            res["[barriers] no_sync"] = timed_b()
        res["[barriers] bucket walk only, wire optimizer (again)"] = timed_b()
        res["[no barriers] bucket walk only, wire optimizer"] = timed(opt, 10)
        with dp.no_sync():
            res["[no barriers] no_sync"] = timed(opt, 10)
    model._dp = None
    res["no DataParallel (again)"] = timed(opt)
    for k, v in res.items():
        print(f"rank {rank}: {v:8.3f} ms  {k}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
