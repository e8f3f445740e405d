"""One extraction step (batch 256, STonKGs-150k shape) after W warm-ups: the short command ncu wraps.

    python tools/profile_step.py [--warmup W] [--steps K] [--batch B] [--layers L] [--train]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from stonkgs_b200 import synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--train", action="store_true")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev, a.layers)
    b = {k: v.to(dev) for k, v in synthetic.make_batch(a.batch, bench.N_KG, seed=3, with_labels=a.train).items()}
    opt = None
    if a.train:   # the bench's pre-training step: train() with dropout, clip + AdamW
        from stonkgs_b200.optim import FusedAdamW
        model.train()
        opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
    from stonkgs_b200 import ops
    for i in range(a.warmup + a.steps):
        l0 = ops.launch_count()
        if a.train:
            opt.zero_grad()
            model(**b)[0].backward()
            opt.step()
        else:
            model.embed(b["input_ids"], b["attention_mask"], b["token_type_ids"])
        per_step = ops.launch_count() - l0
    torch.cuda.synchronize()
    print(f"profile_step done: {per_step} libstk launches per step")


if __name__ == "__main__":
    main()
