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
* With ``FusedAdamW`` attached (it attaches itself when it finds ``model._dp``) there is no unpack pass at all: the
  optimizer reads the summed bf16 wire buffer directly, with 1 / world folded into its clip coefficient — DDP's
  "copy back + divide" fused into the one pass that reads every gradient anyway (about 2 GB of HBM traffic per step
  less at N_kg = 175 003).  ``param.grad`` then holds the rank-LOCAL fp32 gradient; ``materialize_grads()`` writes the
  averaged one back for code that wants to look at it.
* The collective shares the GPU with persistent one-CTA-per-SM GEMM / attention kernels whose CTAs are scheduled
  statically.  When an NCCL CTA holds an SM at such a kernel's launch, the CTA meant for that SM starts only when the
  others finish, and the kernel takes about twice as long: on 2 x B200 the exposed all-reduce time was 0.41-0.47 x the
  collective's own duration whatever ``NCCL_MAX_CTAS`` was (2 / 8 / 32 CTAs: 4.7 / 1.9 / 0.8 ms exposed of 11.4 / 4.1 /
  1.9 ms).  So the persistent kernels leave ``sm_reserve`` SMs free while buckets are in flight (``stk_set_sm_reserve``)
  and NCCL is held to the same number of CTAs: with 4 (bench.py's setting) the collective takes longer (6.6 ms on two
  GPUs) but hides completely behind backward; what remains is the 4 / 148 of the backward GEMMs' throughput and the
  tail bucket (0.6 ms exposed instead of 0.9 with NCCL's default, 8 or 16 reserved SMs cost more than they save).

Gradients are averaged (DDP semantics).  Each cross-entropy is a mean over the *local* labelled rows,
exactly like the reference under DDP.
"""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional

import torch
import torch.distributed as dist


# bring-up only (STK_DP_DEBUG_SKIP_AR=1): run the bucket walk and the pack kernels but not the collective, to separate the
# cost of the orchestration from the cost of the all-reduce itself (gradients are then NOT averaged)
_SKIP_COLLECTIVE = os.environ.get("STK_DP_DEBUG_SKIP_AR", "0") == "1"
_SKIP_PACK = os.environ.get("STK_DP_DEBUG_SKIP_PACK", "0") == "1"


class Bucket:
    __slots__ = ("start", "end", "names", "pending", "work", "event")

    def __init__(self, start):
        self.start, self.end, self.names = start, start, []
        self.pending, self.work, self.event = 0, None, None


def plan_buckets(entries, offsets, bucket_elems: int, tail_elems: int = 0) -> List[Bucket]:
    """Greedy contiguous buckets over the flat buffer in production order.  A segment larger than the
    target gets a bucket of its own (so the entity-decoder gradient ships as soon as it is complete).
    ``tail_elems`` > 0: the LAST bucket — the only one whose all-reduce cannot hide behind backward — is cut short: it
    takes the trailing segments that fit into ``tail_elems`` (at least one segment)."""
    n_tail = 0
    if tail_elems > 0 and len(entries) > 1:
        acc = 0
        for name in reversed(entries):
            padded = (offsets[name][1] + 7) // 8 * 8
            if n_tail > 0 and acc + padded > tail_elems:
                break
            acc += padded
            n_tail += 1
        n_tail = min(n_tail, len(entries) - 1)
    head = entries[: len(entries) - n_tail]
    buckets: List[Bucket] = []
    cur: Optional[Bucket] = None
    for i, name in enumerate(entries):
        if n_tail and i == len(head):
            cur = None   # the tail bucket starts here
        off, n, _ = offsets[name]
        padded = (n + 7) // 8 * 8   # training.GradBuffer.PAD
        in_tail = bool(n_tail) and i >= len(head)
        if cur is None or (not in_tail and (cur.end - cur.start) + padded > bucket_elems and cur.names):
            cur = Bucket(off)
            buckets.append(cur)
        cur.names.append(name)
        cur.end = off + padded
    return buckets


class DataParallel:
    def __init__(self, model, process_group=None, bucket_mb: Optional[float] = None, wire_dtype=torch.bfloat16,
                 overlap: Optional[bool] = None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (one process per GPU)")
        self.model = model
        self.group = process_group
        self.world = dist.get_world_size(process_group)
        if bucket_mb is None:   # fp32 megabytes of gradient per bucket (STK_DP_BUCKET_MB: A/B runs)
            bucket_mb = float(os.environ.get("STK_DP_BUCKET_MB", "64"))
        self.bucket_elems = int(bucket_mb * (1 << 20) / 4)
        # the last bucket's all-reduce is exposed by construction: keep it to the first encoder layer + embeddings
        self.tail_elems = int(float(os.environ.get("STK_DP_TAIL_MB", "32")) * (1 << 20) / 4)
        self.wire_dtype = wire_dtype
        # overlap=False: the same buckets are reduced one after the other once backward has been enqueued (the
        # persistent GEMMs then never share the SMs with a collective kernel); STK_DP_OVERLAP=0 selects it for A/B runs
        self.overlap = (os.environ.get("STK_DP_OVERLAP", "1") != "0") if overlap is None else bool(overlap)
        self.buckets: Optional[List[Bucket]] = None
        self._name_to_bucket = {}
        self._sync = True
        self._stream = None
        self._wire = None
        # SMs the persistent GEMM / attention kernels leave free while buckets are in flight, so that the collective's
        # CTAs always find room.  Only meaningful together with NCCL_MAX_CTAS (read by NCCL when the communicator is
        # created): by default it follows that variable when it is set to a small number, else nothing is reserved.
        ncc = os.environ.get("NCCL_MAX_CTAS", "")
        follow = int(ncc) if ncc.isdigit() and 0 < int(ncc) <= 16 else 0
        self.sm_reserve = int(os.environ.get("STK_DP_SM_RESERVE", follow))
        self.defer_unpack = False     # FusedAdamW attached: it consumes the wire buffer, nothing is unpacked
        self.wire_valid = False       # the wire buffer holds the all-reduced gradient of the last backward
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
            self.buckets = plan_buckets(gb.entries, gb.offsets, self.bucket_elems, self.tail_elems)
            for b in self.buckets:
                for n in b.names:
                    self._name_to_bucket[n] = b
            if gb.flat.is_cuda:
                self._stream = torch.cuda.Stream(device=gb.flat.device)
        if gb.flat.is_cuda and (self._wire is None or self._wire.numel() != gb.flat.numel()):
            # one persistent wire buffer over the whole flat gradient: bucket k travels in its own slice, so buckets in
            # flight never alias and the optimizer can read the reduced gradient in place
            self._wire = torch.zeros(gb.flat.numel(), dtype=self.wire_dtype, device=gb.flat.device)
        self._gb = gb
        self.wire_valid = False
        if self.sm_reserve and gb.flat.is_cuda and self._sync and self.overlap:
            from . import _lib
            _lib.load().stk_set_sm_reserve(gb.flat.device.index, self.sm_reserve)
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
                wire = self._wire[b.start:b.end]
                if not _SKIP_PACK:
                    ops.cast_bf16(flat, out=wire)                # pack: fp32 -> bf16 on the wire
                if not _SKIP_COLLECTIVE:
                    dist.all_reduce(wire, op=dist.ReduceOp.SUM, group=self.group)
                if not self.defer_unpack:
                    ops.unpack_scale(wire, flat, 1.0 / self.world)   # unpack: mean, back to fp32
            else:
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.mul_(1.0 / self.world)
            b.event = torch.cuda.Event()
            b.event.record(self._stream)

    def finish(self, gb):
        if self.sm_reserve and gb.flat.is_cuda:
            from . import _lib
            _lib.load().stk_set_sm_reserve(gb.flat.device.index, 0)
        if not self._sync:
            return
        for b in self.buckets:
            if b.pending != 0:
                raise RuntimeError(f"gradient bucket {b.names[0]}.. was never completed by backward")
            if not self.overlap:
                self._reduce(b)
            if b.event is not None:
                torch.cuda.current_stream(gb.flat.device).wait_event(b.event)
        self.wire_valid = bool(self.defer_unpack and gb.flat.is_cuda and self.wire_dtype == torch.bfloat16)

    def attach_optimizer(self, opt) -> bool:
        """Called by FusedAdamW: from now on the optimizer reads the reduced bf16 wire buffer itself (scaled by 1 / world)
        and no unpack pass runs.  Returns False (nothing changes) for an fp32 wire."""
        if self.wire_dtype != torch.bfloat16:
            return False
        self.defer_unpack = True
        return True

    def wire_buffer(self, gb):
        if self._wire is None or self._wire.numel() != gb.flat.numel():
            self._wire = torch.zeros(gb.flat.numel(), dtype=self.wire_dtype, device=gb.flat.device)
        return self._wire

    def materialize_grads(self):
        """Write the averaged gradient of the last synchronised backward into ``param.grad`` (the flat fp32 buffer).
        Only needed with an attached FusedAdamW, and only by code that inspects gradients."""
        if self.wire_valid:
            from . import ops
            ops.unpack_scale(self._wire, self._gb.flat, 1.0 / self.world)
            self.wire_valid = False   # the fp32 buffer is authoritative again (the optimizer falls back to it)
