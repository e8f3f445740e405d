"""GPU busy time vs wall time of the bench steps (torch.profiler / Kineto): shows launch gaps and host syncs.

    python tools/gpu_busy.py [--train] [--steps K]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
from stonkgs_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev, 12)
    B = 64 if a.train else 256
    b = {k: v.to(dev) for k, v in synthetic.make_batch(B, bench.N_KG, seed=3, with_labels=a.train).items()}
    opt = None
    if a.train:
        from stonkgs_b200.optim import FusedAdamW
        model.train()
        opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)

    def step():
        if a.train:
            opt.zero_grad()
            model(**b)[0].backward()
            opt.step()
        else:
            model.embed(b["input_ids"], b["attention_mask"], b["token_type_ids"])

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            step()
        e1.record()
        torch.cuda.synchronize()
    wall = e0.elapsed_time(e1)
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    spans = sorted((e.time_range.start, e.time_range.end) for e in evs)
    busy, cur_s, cur_e = 0.0, None, None
    for s, e in spans:                       # union of kernel intervals
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        busy += cur_e - cur_s
    gaps = []
    last = None
    for s, e in spans:
        if last is not None and s > last:
            gaps.append(s - last)
        last = e if last is None else max(last, e)
    gaps.sort(reverse=True)
    print(f"steps {a.steps}: wall {wall:.2f} ms, GPU busy {busy / 1000:.2f} ms ({busy / 10 / wall:.1f} %), kernels {len(spans)}")
    print("largest gaps (us):", [round(g, 1) for g in gaps[:12]], "sum of gaps > 5 us:", round(sum(g for g in gaps if g > 5) / 1000, 2), "ms")


if __name__ == "__main__":
    main()
