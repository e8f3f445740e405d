"""Data-parallel pre-training: bucketed gradient all-reduce overlapped with the hand-written backward.

Replaces what the reference gets implicitly from ``accelerate`` / HF ``Trainer`` wrapping the model
in ``torch.nn.parallel.DistributedDataParallel`` (stonkgs_pretraining.py:147-168,215-223; SURVEY
§2.1).  One process per GPU; ``torch.distributed`` (NCCL over NVLink / NVSwitch) is the plumbing.

* Gradients live in ONE flat fp32 buffer laid out in the order backward produces them
  (``training.GradBuffer``: entity decoder first, embeddings last); only the 206 live tensors are in
  it, so there is no ``find_unused_parameters`` graph walk (the six dead tensors of the reference head
  never get a gradient).
* The buffer is cut into contiguous buckets.  ``backward`` reports each finished segment
  (``on_ready``); when a bucket is complete an event is recorded on the compute stream and, on a side
  stream: pack fp32 -> bf16 (libstk cast kernel) -> ``all_reduce(SUM)`` -> unpack * 1/world (libstk
  kernel) back into the flat buffer.  The largest bucket (entity decoder, > half of all gradient
  bytes) becomes ready first, so its transfer hides behind the whole trunk backward.
* ``finish`` makes the compute stream wait for the last bucket; ``no_sync()`` skips the collective for
  gradient-accumulation micro-steps (the sum is reduced once, by the last micro-step).

Gradients are averaged (DDP semantics).  Each cross-entropy is a mean over the *local* labelled rows,
exactly like the reference under DDP.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional

import torch
import torch.distributed as dist


class Bucket:
    __slots__ = ("start", "end", "names", "pending", "work", "event")

    def __init__(self, start):
        self.start, self.end, self.names = start, start, []
        self.pending, self.work, self.event = 0, None, None


def plan_buckets(entries, offsets, bucket_elems: int) -> List[Bucket]:
    """Greedy contiguous buckets over the flat buffer in production order.  A segment larger than the
    target gets a bucket of its own (so the entity-decoder gradient ships as soon as it is complete)."""
    buckets: List[Bucket] = []
    cur: Optional[Bucket] = None
    for name in entries:
        off, n, _ = offsets[name]
        padded = (n + 3) // 4 * 4
        if cur is None or (cur.end - cur.start) + padded > bucket_elems and cur.names:
            cur = Bucket(off)
            buckets.append(cur)
        cur.names.append(name)
        cur.end = off + padded
    return buckets


class DataParallel:
    def __init__(self, model, process_group=None, bucket_mb: float = 64.0, wire_dtype=torch.bfloat16,
                 overlap: Optional[bool] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        self.wire_dtype = wire_dtype
        # overlap=False: the same buckets are reduced one after the other once backward has been enqueued (the
        # persistent GEMMs then never share the SMs with a collective kernel); STK_DP_OVERLAP=0 selects it for A/B runs
        self.overlap = (os.environ.get("STK_DP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.buckets: Optional[List[Bucket]] = None
        self._name_to_bucket = {}
        self._sync = True
        self._stream = None
        self._wire = None
        model._dp = self
        self.broadcast_parameters()

    def broadcast_parameters(self):
        """Rank 0's weights everywhere (DDP does the same at construction)."""
        for t in list(self.model.parameters()) + list(self.model.buffers()):
            dist.broadcast(t.data, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                           group=self.group)
        # writes through .data do not bump Parameter._version, which is what the model's staleness check of its
        # device-side state looks at: drop that state (bf16 GEMM copies, fused q|k|v bias, the three LM-backbone rows
        # of the KG table) so that every rank rebuilds it from the weights it just received
        if hasattr(self.model, "_dev_state"):
            self.model._dev_state = None
            self.model._special_rows_version = None

    @contextlib.contextmanager
    def no_sync(self):
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    # ---- hooks called by training._PretrainStep.backward ------------------------------------------
    def begin(self, gb):
        if self.buckets is None:
            self.buckets = plan_buckets(gb.entries, gb.offsets, self.bucket_elems)
            for b in self.buckets:
                for n in b.names:
                    self._name_to_bucket[n] = b
            if gb.flat.is_cuda:
                self._stream = torch.cuda.Stream(device=gb.flat.device)
                biggest = max(b.end - b.start for b in self.buckets)
                self._wire = torch.empty(biggest, dtype=self.wire_dtype, device=gb.flat.device)
        self._gb = gb
        for b in self.buckets:
            b.pending, b.work, b.event = len(b.names), None, None

    def on_ready(self, name: str):
        b = self._name_to_bucket[name]
        b.pending -= 1
        if b.pending == 0 and self._sync and self.overlap:
            self._reduce(b)

    def _reduce(self, b: Bucket):
        flat = self._gb.flat[b.start:b.end]
        if not flat.is_cuda:
            # host-side logic tests (gloo): same bucket walk, plain fp32 all-reduce
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            flat.mul_(1.0 / self.world)
            return
        from . import ops
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(flat.device))
        with torch.cuda.stream(self._stream):
            self._stream.wait_event(ready)
            if self.wire_dtype == torch.bfloat16:
                wire = self._wire[: b.end - b.start]
                ops.cast_bf16(flat, out=wire)                    # pack: fp32 -> bf16 on the wire
                dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
                ops.unpack_scale(wire, flat, 1.0 / self.world)   # unpack: mean, back to fp32
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.mul_(1.0 / self.world)
            b.event = torch.cuda.Event()
            b.event.record(self._stream)

    def finish(self, gb):
        if not self._sync:
            return
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError(f"gradient bucket {b.names[0]}.. was never completed by backward")
            if not self.overlap:
                self._reduce(b)
            if b.event is not None:
                torch.cuda.current_stream(gb.flat.device).wait_event(b.event)
