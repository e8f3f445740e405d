"""bench.py — headline benchmark of the STonKGs hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--workload extract|pretrain] [--impl reference]

Default workload (BASELINE.json configs[1]): ``get_stonkgs_embeddings``-style extraction,
STonKGs-150k shape (12+12 BERT-base layers, 256 text + 256 KG tokens, N_kg = 175 003), batch 256 per
GPU, bf16 tensor-core compute; metric = text-triple pairs / second.  One "step" = one batch of 256
pairs through LM backbone -> KG lookup -> joint encoder -> pooler.  N > 1 (torchrun): every rank
embeds its own shard of the pairs, no data-path collective (weak scaling).

The JSON line carries
  value     pairs/s with the step's inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same through host buffers: pinned int64 ids -> H2D -> forward -> D2H pooled [256,768]
  roofline  the dominant kernel's achieved TFLOP/s (algorithmic FLOPs / CUDA-event time inside a
            profiled step) against the measured cuBLAS bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (fp32 torch CPU restatement of the reference) on the host cores,
            bounded sample (N=1, rank 0 only)
``--impl reference`` times the CPU oracle port alone (the reference is pure Python on HF BERT and
cannot be pip-installed offline: its import needs pystow/indra/pybel + network).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("HF_HUB_OFFLINE", "1")
os.environ.setdefault("TRANSFORMERS_OFFLINE", "1")

N_KG = 175003
BATCH = 256            # pairs per GPU per step (extraction)
TRAIN_BATCH = 64       # pairs per GPU per step (pre-training; global 512 at 8 GPUs)
GFLOP_PER_PAIR_EXTRACT = 142.54
GFLOP_PER_PAIR_TRAIN = 371.8


# --------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_burst": d.get("bf16_tflops"), "bf16_sustained": d.get("bf16_tflops_sustained"),
                "hbm": d.get("hbm_gbs"), "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed regions."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pw.append(0.0)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing at least 60 % of the highest power seen (an idle GPU sits at its maximum clock,
        # a power-capped busy one well below it, so the clock value itself cannot tell the two apart)
        top = max(pw) if pw else 0.0
        busy = sorted(c for c, w in zip(sm, pw) if top <= 0.0 or w >= 0.6 * top)
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons), "samples": len(sm),
                "samples_under_load": len(busy), "power_w_max": top or None}


def dist_env():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return world, rank, local


# --------------------------------------------------------------------------------------------------
def build_model(device, layers=12, n_kg=N_KG, seed=0):
    import torch
    from transformers import BertConfig
    import numpy as np
    from stonkgs_b200.model import STonKGsForPreTraining
    torch.manual_seed(seed)
    rows = np.random.default_rng(seed).standard_normal((n_kg, 768)).astype(np.float32)  # synthetic node2vec file
    model = STonKGsForPreTraining(None, BertConfig(vocab_size=28996, num_hidden_layers=layers), rows)
    # non-trivial biases / LayerNorm gains so that no kernel sees an all-zero vector
    with torch.no_grad():
        g = torch.Generator().manual_seed(seed + 1)
        for n, p in model.named_parameters():
            if n.endswith(".bias") and p.dim() == 1:
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
    return model.eval().to(device)


def cpu_port_throughput(n_pairs: int, seed=0):
    """Oracle port (fp32 torch restatement of the reference forward) on all host cores."""
    import torch
    from oracle import stonkgs_oracle as orc, weights
    from stonkgs_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_kg = 3001  # the extraction forward never touches the decoders; a small table keeps set-up short
    sd = weights.make_state_dict(n_kg, 12, seed)
    table = torch.from_numpy(weights.make_kg_table(n_kg, seed))
    table = torch.cat([table, torch.zeros(3, 768)])
    bs = 8
    batch = synthetic.make_batch(bs, n_kg, seed=1, with_labels=False)
    with torch.no_grad():
        orc.forward(sd, table, **batch)  # warm-up
        t0 = time.perf_counter()
        done = 0
        while done < n_pairs:
            orc.forward(sd, table, **batch)
            done += bs
        dt = time.perf_counter() - t0
    return done / dt, cores, f"{done} pairs as batches of {bs}, 12+12 layers, fp32, eval forward (extraction path)"


def run_reference(args):
    world, rank, _ = dist_env()
    if rank != 0:
        return
    t0 = time.perf_counter()
    per_step = 8
    import torch
    from oracle import stonkgs_oracle as orc, weights
    from stonkgs_b200 import synthetic
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    n_kg = 3001
    sd = weights.make_state_dict(n_kg, 12, 0)
    table = torch.cat([torch.from_numpy(weights.make_kg_table(n_kg, 0)), torch.zeros(3, 768)])
    batch = synthetic.make_batch(per_step, n_kg, seed=1, with_labels=False)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            orc.forward(sd, table, **batch)
        steps = min(args.steps, 6)   # bounded: a CPU step takes seconds
        t1 = time.perf_counter()
        for _ in range(steps):
            orc.forward(sd, table, **batch)
        dt = time.perf_counter() - t1
    v = steps * per_step / dt
    sample = f"{steps} steps x {per_step} pairs (bounded sample of the batch-256 workload), fp32 CPU"
    print(json.dumps({
        "impl": "reference", "metric": "text-triple pairs/sec (embedding extraction)", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1000 * dt / steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "get_stonkgs_embeddings-style extraction, STonKGs-150k shape, CPU oracle port of the "
                               "reference forward (the reference package itself is not importable offline)"},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0}))


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="extract", choices=["extract", "pretrain"])
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-dropout", action="store_true", help="pretrain workload: switch the train()-mode dropout off")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from stonkgs_b200 import ops, synthetic

    world, rank, local = dist_env()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    train = args.workload == "pretrain"
    B = args.batch or (TRAIN_BATCH if train else BATCH)
    model = build_model(dev, args.layers)
    opt = None
    if train:
        # train() like the reference's Trainer loop: hidden / attention-probability dropout (p = 0.1) is ON, with
        # counter-based masks regenerated in the backward pass; --no-dropout times the eval-numerics step
        model.train()
        model.stk_dropout = not args.no_dropout
        if world > 1:
            from stonkgs_b200.dp import DataParallel
            DataParallel(model, dist.group.WORLD)
        from stonkgs_b200.optim import FusedAdamW
        # HF Trainer defaults of the reference driver: AdamW lr 1e-4, wd 0, max_grad_norm 1.0
        opt = FusedAdamW(model, lr=1e-4, weight_decay=0.0, max_grad_norm=1.0)
    # distinct batches per step so that no step re-reads the previous step's inputs from L2
    n_batches = 4
    host = [synthetic.make_batch(B, N_KG, seed=100 + rank * 17 + i, with_labels=train) for i in range(n_batches)]
    host = [{k: v.pin_memory() for k, v in b.items()} for b in host]
    resident = [{k: v.to(dev) for k, v in b.items()} for b in host]
    pooled_host = torch.empty((B, 768), dtype=torch.float32).pin_memory()
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step_resident(i):
        b = resident[i % n_batches]
        if train:
            opt.zero_grad()
            loss = model(**b)[0]
            loss.backward()
            opt.step()
            return loss
        return model.embed(b["input_ids"], b["attention_mask"], b["token_type_ids"])

    def step_e2e(i):
        b = {k: v.to(dev, non_blocking=True) for k, v in host[i % n_batches].items()}
        if train:
            opt.zero_grad()
            loss = model(**b)[0]
            loss.backward()
            opt.step()
            loss_host.copy_(loss.detach(), non_blocking=True)
        else:
            pooled_host.copy_(model.embed(b["input_ids"], b["attention_mask"], b["token_type_ids"]), non_blocking=True)
        torch.cuda.synchronize(dev)   # the caller reads the result every step

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ops.launch_count()
    ms = timed(step_resident, args.steps)
    launches = ops.launch_count() - launches0
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel: one profiled step, CUDA events around every GEMM/attention launch
    prof = ops.LaunchProfiler()
    ops.set_profiler(prof)
    step_resident(0)
    agg = prof.summary()
    ops.set_profiler(None)
    peaks = measured_peaks()
    total_ms = sum(a["ms"] for a in agg.values())
    gemm = {k: v for k, v in agg.items() if k.startswith("gemm")}
    dom_name = max(gemm, key=lambda k: gemm[k]["ms"]) if gemm else None   # dominant kernel instantiation
    dom = gemm[dom_name] if dom_name else None
    g_ms = sum(a["ms"] for a in gemm.values())
    g_fl = sum(a["work"] for a in gemm.values())
    achieved = dom["work"] / dom["ms"] / 1e9 if dom else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if dom_name and os.path.exists(tpath):
        t = json.load(open(tpath)).get(dom_name)
        if t:   # ncu-measured DRAM bytes of one captured launch, scaled to this run's average launch
            traffic = t["dram_bytes"] * (dom["work"] / dom["launches"]) / t["flops"]
    names = {"gemm_a0b0_epi0": "fused QKV projection", "gemm_a0b0_epi1": "FFN1 + erf-GELU epilogue",
             "gemm_a0b0_epi3": "Wo / FFN2 + residual epilogue", "gemm_a0b0_epi10": "Wo / FFN2 + bias + residual + LayerNorm epilogue (3-CTA cluster)", "gemm_a1b1_epi6": "wgrad (split-K reduce-add)",
             "gemm_a0b1_epi3": "dgrad + residual", "gemm_a0b1_epi5": "dgrad * GELU'", "gemm_a0b1_epi12": "dgrad * saved GELU'",
             "gemm_a0b0_epi11": "FFN1 + erf-GELU epilogue saving GELU'"}
    roofline = {
        "bound": "tensor",
        "kernel": f"stk::gemm_kernel<{dom_name}> ({names.get(dom_name, 'tcgen05 GEMM')})" if dom_name else None,
        "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": (achieved / peaks["bf16_sustained"]) if achieved else None,
        "peak_source": f"{peaks['source']} cuBLAS bf16 sustained (kernel timed inside a long step)",
        "traffic": traffic,
        "algorithmic_flops_per_launch": (dom["work"] / dom["launches"]) if dom else None,
        "launches": dom["launches"] if dom else None, "ms_per_launch": (dom["ms"] / dom["launches"]) if dom else None,
        "share_of_profiled_step": (dom["ms"] / total_ms) if dom and total_ms else None,
        "gemm_family": {"achieved": g_fl / g_ms / 1e9 if g_ms else None, "share_of_profiled_step": g_ms / total_ms if total_ms else None},
        "per_kernel": {k: {"ms": round(v["ms"], 3), "tflops": round(v["work"] / v["ms"] / 1e9, 1), "launches": v["launches"]}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])},
    }

    pairs = world * B * args.steps
    value = pairs / (ms / 1000)
    e2e_value = pairs / (ms_e2e / 1000)
    h2d = sum(v.numel() * v.element_size() for v in host[0].values())
    d2h = 4 if train else B * 768 * 4
    gflop = GFLOP_PER_PAIR_TRAIN if train else GFLOP_PER_PAIR_EXTRACT
    out = {
        "metric": "text-triple pairs/sec (" + ("pretrain step" if train else "embedding extraction") + ")",
        "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": ("STonKGs-150k-shape pretraining step (fwd + bwd + MLM/ELM/NSP losses + DP grad allreduce + clip + AdamW; "
                                + ("dropout off" if args.no_dropout else "train() mode with dropout 0.1") + ")"
                                if train else "get_stonkgs_embeddings-style extraction, STonKGs-150k shape"),
                   "batch_per_gpu": B, "global_batch": B * world, "seq_len": "256 text + 256 KG", "layers": f"{args.layers}+{args.layers}",
                   "kg_vocab": N_KG, "parallelism": f"dp{world}" if train else f"batch-sharded x{world}, no comms",
                   "l2_policy": "per-step working set (>1 GB activations) exceeds the 126 MB L2; 4 rotating input batches"},
        "model_tflops_per_gpu": value / world * gflop / 1000,
        "frac_of_bf16_sustained_peak": value / world * gflop / 1000 / peaks["bf16_sustained"],
        "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample = cpu_port_throughput(24)
        out["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
