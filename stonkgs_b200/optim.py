"""Fused clip + AdamW for the live parameters of ``STonKGsForPreTraining``.

The reference trains with HF ``Trainer`` defaults (``stonkgs_pretraining.py:171-193``): AdamW
(lr 1e-4, betas 0.9/0.999, eps 1e-8, weight decay 0), ``max_grad_norm=1.0``, linear LR decay.  This
optimizer performs the same update as ``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.AdamW.step``
as two libstk.so launches over the flat gradient buffer (global norm, then one multi-tensor AdamW
pass) and refreshes the bf16 GEMM copies of the weights in the same pass, so no cast pass follows.

It is a ``torch.optim.Optimizer`` (single param group; ``lr`` is read from the group each step, so HF
LR schedulers work) and can be handed to ``Trainer(optimizers=(opt, scheduler))``.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import AdamSeg, StkError, check

CHUNK = 65536


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=1.0):
        self.model = model
        gb = model.grad_buffer()
        params = [p for p, _ in gb.param_views]
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      max_grad_norm=max_grad_norm))
        self._gb = gb
        dev = gb.flat.device
        self.exp_avg = torch.zeros_like(gb.flat)
        self.exp_avg_sq = torch.zeros_like(gb.flat)
        self._sumsq = torch.zeros(1, dtype=torch.float32, device=dev)
        self._step = 0
        # data parallel: consume the all-reduced bf16 wire buffer directly (no unpack pass), see dp.DataParallel
        self._dp = getattr(model, "_dp", None)
        self._use_wire = bool(self._dp is not None and gb.flat.is_cuda and self._dp.attach_optimizer(self))
        self._build_tables(dev)

    # ------------------------------------------------------------------------------------------
    def _bf16_destinations(self):
        """name of flat-buffer segment -> (bf16 GEMM copy, fused fp32 copy) of the device-side weights."""
        st = self.model._device_state(need_heads=True)
        bert, heads = st["bert"], st["heads"]
        w16 = {"w_ent": heads.w_ent, "w_text": heads.w_text, "t_w": heads.wt, "pool_w": bert.wp}
        p32 = {}
        for i, lw in enumerate(bert.layers):
            w16[f"l{i}.w2"], w16[f"l{i}.w1"], w16[f"l{i}.wo"], w16[f"l{i}.wqkv"] = lw.w2, lw.w1, lw.wo, lw.wqkv
            p32[f"l{i}.bqkv"] = lw.bqkv
        return w16, p32

    def _build_tables(self, dev):
        gb = self._gb
        w16, p32 = self._bf16_destinations()
        view_to_seg = {}
        for name, v in gb.views.items():
            view_to_seg[v.data_ptr()] = name
        segs, chunk_seg, chunk_off = [], [], []
        base = gb.flat.data_ptr()
        wire = self._dp.wire_buffer(gb) if self._use_wire else None
        keep = []
        # one AdamSeg per parameter; the fused q|k|v gradient block maps onto three parameters
        for p, gv in gb.param_views:
            if not p.data.is_contiguous():
                raise StkError("FusedAdamW needs contiguous parameters")
            off = (gv.data_ptr() - base) // 4
            n = p.numel()
            # which named segment holds this view, and at which element offset inside it
            seg_name, inner = None, 0
            for name, (soff, sn, _) in gb.offsets.items():
                if soff <= off < soff + sn:
                    seg_name, inner = name, off - soff
                    break
            s = AdamSeg()
            s.p = p.data.data_ptr()
            s.g = gv.data_ptr()
            s.m = self.exp_avg.data_ptr() + off * 4
            s.v = self.exp_avg_sq.data_ptr() + off * 4
            s.w16 = (w16[seg_name].data_ptr() + inner * 2) if seg_name in w16 else None
            s.p32_copy = (p32[seg_name].data_ptr() + inner * 4) if seg_name in p32 else None
            s.n = n
            s.g16 = None
            idx = len(segs)
            segs.append(s)
            for c in range(0, n, CHUNK):
                chunk_seg.append(idx)
                chunk_off.append(c)
        arr = (AdamSeg * len(segs))(*segs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self._segs_dev = host.to(dev)
        self._segs_wire_dev = None
        if wire is not None:   # the same segments reading the bf16 wire buffer instead of the fp32 gradient
            for sg in segs:
                sg.g16 = wire.data_ptr() + ((sg.g - base) // 4) * 2
            arr = (AdamSeg * len(segs))(*segs)
            self._segs_wire_dev = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(dev)
            self._wire = wire
        self._chunk_seg = torch.tensor(chunk_seg, dtype=torch.int32, device=dev)
        self._chunk_off = torch.tensor(chunk_off, dtype=torch.int64, device=dev)
        self._n_chunks = len(chunk_seg)
        self._runs = gb.trainable_runs()
        self._weights_state = self.model._dev_state   # tables point into these buffers

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        gb = self._gb
        if self.model.grad_buffer() is not gb:
            raise StkError("the model's gradient buffer was rebuilt (moved to another device, or requires_grad of a "
                           "parameter changed) after FusedAdamW was constructed: build a new optimizer")
        self.model._raise_on_bad_ids()   # deferred id / label range check of this step's forward (backward is enqueued)
        if self.model._dev_state is not self._weights_state or self.model._dev_state is None:
            self._build_tables(gb.flat.device)   # the model was moved / its device state rebuilt
        if any(p.grad is None or p.grad.data_ptr() != v.data_ptr() for p, v in gb.param_views):
            raise StkError("FusedAdamW.step: gradients are not in the model's flat gradient buffer "
                           "(run loss.backward() of STonKGsForPreTraining first)")
        g = self.param_groups[0]
        self._step += 1
        b1, b2 = g["betas"]
        dev = gb.flat.device
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib = _lib.load()
        clip = g["max_grad_norm"] is not None and g["max_grad_norm"] > 0
        # data parallel with the collective done for this backward: gradients = wire buffer (sum over ranks) / world
        from_wire = self._segs_wire_dev is not None and self._dp.wire_valid and self._dp._wire is self._wire
        gscale = 1.0 / self._dp.world if from_wire else 1.0
        if clip:
            self._sumsq.zero_()
            for a, b in self._runs:   # one range unless some live parameters are frozen
                if from_wire:
                    check(lib.stk_sumsq_bf16(dev.index, stream, ctypes.c_void_p(self._wire.data_ptr() + 2 * a), b - a,
                                             gscale, ctypes.c_void_p(self._sumsq.data_ptr())), "stk_sumsq_bf16")
                else:
                    check(lib.stk_sumsq(dev.index, stream, ctypes.c_void_p(gb.flat.data_ptr() + 4 * a), b - a,
                                        ctypes.c_void_p(self._sumsq.data_ptr())), "stk_sumsq")
        segs_dev = self._segs_wire_dev if from_wire else self._segs_dev
        check(lib.stk_adamw_step(dev.index, stream, ctypes.c_void_p(segs_dev.data_ptr()),
                                 ctypes.c_void_p(self._chunk_seg.data_ptr()), ctypes.c_void_p(self._chunk_off.data_ptr()),
                                 self._n_chunks, float(g["lr"]), float(b1), float(b2), float(g["eps"]),
                                 float(g["weight_decay"]), 1.0 - b1 ** self._step, 1.0 - b2 ** self._step,
                                 ctypes.c_void_p(self._sumsq.data_ptr()) if clip else None,
                                 float(g["max_grad_norm"] or 0.0), gscale), "stk_adamw_step")
        return loss

    def grad_norm(self) -> torch.Tensor:
        """Global gradient norm seen by the last step (device scalar, no sync)."""
        return self._sumsq.sqrt()

    def zero_grad(self, set_to_none: bool = True):
        for p, _ in self._gb.param_views:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    def state_dict(self):
        return {"step": self._step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in self.param_groups[0].items() if k != "params"}]}

    def load_state_dict(self, sd):
        self._step = int(sd["step"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.param_groups[0].update(sd["param_groups"][0])
