"""torch-tensor front ends of the C ABI (include/stk.h).

PyTorch is used for device memory and streams only; every function here launches hand-written
sm_100a kernels from libstk.so on the current CUDA stream and fails loudly otherwise.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import (EPI_BIAS, EPI_BIAS_DROP_RESID_LN, EPI_BIAS_GELU, EPI_BIAS_GELU_SAVE, EPI_BIAS_RESID, EPI_BIAS_RESID_LN,
                   EPI_BIAS_TANH_F32, EPI_BIAS_GELU_SAVE_GRAD, EPI_CE_DLOGIT, EPI_CE_STATS, EPI_DGELU, EPI_F32, EPI_F32_ADD,
                   EPI_MUL, WS_ATTN_BWD, WS_LINEAR_CE_BWD, WS_LINEAR_CE_FWD, GemmEpilogue, check)

H = 768
HEADS = 12

_F32_EPIS = (EPI_F32, EPI_F32_ADD, EPI_BIAS_TANH_F32)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _ctx(t: torch.Tensor):
    if not t.is_cuda:
        raise _lib.StkError("stonkgs_b200 kernels need CUDA tensors (there is no CPU path)")
    dev = t.device.index if t.device.index is not None else torch.cuda.current_device()
    return dev, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _req(t: torch.Tensor, dtype, name: str):
    if t.dtype != dtype or not t.is_cuda:
        raise _lib.StkError(f"{name}: expected CUDA {dtype}, got {t.dtype} on {t.device}")
    return t


def launch_count() -> int:
    return int(_lib.load().stk_launch_count())


class LaunchProfiler:
    """Optional per-launch CUDA-event timing of the GEMM / attention wrappers (bench.py roofline).
    Events are recorded on the launching stream; nothing is synchronised until ``summary()``."""

    def __init__(self):
        self.records = []  # (name, work, start_event, end_event)

    def span(self, name, work):
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        self.records.append((name, work, e0, e1))
        return e0, e1

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, work, e0, e1 in self.records:
            a = agg.setdefault(name, {"ms": 0.0, "work": 0.0, "launches": 0})
            a["ms"] += e0.elapsed_time(e1)
            a["work"] += work
            a["launches"] += 1
        return agg


_PROFILER: Optional[LaunchProfiler] = None


def set_profiler(p: Optional[LaunchProfiler]) -> None:
    global _PROFILER
    _PROFILER = p


# --------------------------------------------------------------------------------------------------
# embedding stages / LayerNorm
# --------------------------------------------------------------------------------------------------
def embed_text_ln(ids: torch.Tensor, word, pos, type_emb, gamma, beta, err_flag=None) -> torch.Tensor:
    """ids: int64 [B, S] (may be a column slice of a wider tensor). Returns bf16 [B*S, 768]."""
    _req(ids, torch.int64, "ids")
    B, S = ids.shape
    assert ids.stride(1) == 1
    dev, stream = _ctx(ids)
    out = torch.empty((B * S, H), dtype=torch.bfloat16, device=ids.device)
    check(_lib.load().stk_embed_text_ln_fwd(dev, stream, _ptr(ids), ids.stride(0), B, S, _ptr(word), word.shape[0],
                                            _ptr(pos), _ptr(type_emb), _ptr(gamma), _ptr(beta), _ptr(out),
                                            _ptr(err_flag)), "stk_embed_text_ln_fwd")
    return out


class SeqShape:
    """Joint sequence of one pair: ``text_len`` LM tokens + KG tokens = ``seq_len`` ids; activations hold ``seq_pad``
    rows per pair (a multiple of 128 for the attention kernels; rows >= seq_len are zero and masked out as keys).
    STonKGs: (256, 512, 512).  TransE variant (transestonkgs_model.py:44,93): (256, 260, 384)."""
    __slots__ = ("text_len", "seq_len", "seq_pad")

    def __init__(self, text_len: int = 256, seq_len: int = 512, seq_pad: Optional[int] = None):
        self.text_len, self.seq_len = int(text_len), int(seq_len)
        self.seq_pad = int(seq_pad) if seq_pad is not None else (self.seq_len + 127) // 128 * 128
        if not (0 < self.text_len < self.seq_len <= self.seq_pad <= 512 and self.seq_pad % 128 == 0):
            raise _lib.StkError(f"unsupported joint sequence shape {self.text_len}+{self.seq_len - self.text_len} "
                           f"(padded {self.seq_pad}): at most 512 positions")

    @property
    def kg_len(self) -> int:
        return self.seq_len - self.text_len

    def __repr__(self):
        return f"SeqShape({self.text_len}, {self.seq_len}, {self.seq_pad})"


STONKGS_SHAPE = SeqShape(256, 512, 512)


def embed_joint_ln(input_ids, token_type_ids, lm_hidden, kg_table, pos, type_emb, gamma, beta, *, save_stats=False,
                   want_inputs_embeds=False, err_flag=None, shape: SeqShape = STONKGS_SHAPE):
    """Returns (out bf16 [B*seq_pad,768], mean, rstd, inputs_embeds fp32 or None)."""
    _req(input_ids, torch.int64, "input_ids")
    B = input_ids.shape[0]
    SP = shape.seq_pad
    assert input_ids.shape[1] == shape.seq_len and input_ids.is_contiguous()
    assert pos.shape[0] >= shape.seq_len and lm_hidden.numel() == B * shape.text_len * H and lm_hidden.is_contiguous()
    if token_type_ids is not None:
        _req(token_type_ids, torch.int64, "token_type_ids")
        assert token_type_ids.is_contiguous() and token_type_ids.shape == input_ids.shape
    _req(lm_hidden, torch.bfloat16, "lm_hidden")
    _req(kg_table, torch.float32, "kg_table")
    dev, stream = _ctx(input_ids)
    out = torch.empty((B * SP, H), dtype=torch.bfloat16, device=input_ids.device)
    mean = rstd = emb = None
    if save_stats:
        mean = torch.empty(B * SP, dtype=torch.float32, device=input_ids.device)
        rstd = torch.empty_like(mean)
    if want_inputs_embeds:
        emb = torch.empty((B * SP, H), dtype=torch.float32, device=input_ids.device)
    check(_lib.load().stk_embed_joint_ln_fwd_shape(dev, stream, _ptr(input_ids), _ptr(token_type_ids), B, shape.text_len,
                                                   shape.seq_len, SP, _ptr(lm_hidden), _ptr(kg_table), kg_table.shape[0],
                                                   _ptr(pos), _ptr(type_emb), _ptr(gamma), _ptr(beta), _ptr(out),
                                                   _ptr(mean), _ptr(rstd), _ptr(emb), _ptr(err_flag)),
          "stk_embed_joint_ln_fwd")
    return out, mean, rstd, emb


def embed_joint_ln_bwd(input_ids, token_type_ids, lm_hidden, kg_table, pos, type_emb, gamma, mean, rstd, dy, dpos,
                       dtype, dgamma, dbeta, shape: SeqShape = STONKGS_SHAPE):
    B = input_ids.shape[0]
    dev, stream = _ctx(input_ids)
    assert dy.shape[0] == B * shape.seq_pad and dpos.shape[0] >= shape.seq_len
    check(_lib.load().stk_embed_joint_ln_bwd_shape(dev, stream, _ptr(input_ids), _ptr(token_type_ids), B, shape.text_len,
                                                   shape.seq_len, shape.seq_pad, _ptr(lm_hidden), _ptr(kg_table),
                                                   kg_table.shape[0], _ptr(pos), _ptr(type_emb), _ptr(gamma), _ptr(mean),
                                                   _ptr(rstd), _ptr(dy), _ptr(dpos), _ptr(dtype), _ptr(dgamma),
                                                   _ptr(dbeta)), "stk_embed_joint_ln_bwd")


def layernorm(x: torch.Tensor, gamma, beta, save_stats=False, out=None):
    _req(x, torch.bfloat16, "x")
    M = x.shape[0]
    assert x.shape[1] == H and x.is_contiguous()
    dev, stream = _ctx(x)
    y = torch.empty_like(x) if out is None else out
    mean = rstd = None
    if save_stats:
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
    check(_lib.load().stk_layernorm_fwd(dev, stream, _ptr(x), M, _ptr(gamma), _ptr(beta), _ptr(y), _ptr(mean),
                                        _ptr(rstd)), "stk_layernorm_fwd")
    return (y, mean, rstd) if save_stats else y


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, out=None, dbias=None, drop: Optional["Drop"] = None):
    """dgamma/dbeta are accumulated into. Returns dx (bf16); with ``dbias`` / ``drop`` the fused form: the column sums
    of the (masked) result are accumulated into ``dbias`` and, with dropout, ``(dx, dx through the mask)`` is returned."""
    M = x.shape[0]
    dev, stream = _ctx(x)
    dx = torch.empty_like(x) if out is None else out
    if dbias is None and drop is None:
        check(_lib.load().stk_layernorm_bwd(dev, stream, _ptr(dy), _ptr(x), M, _ptr(gamma), _ptr(mean), _ptr(rstd),
                                            _ptr(dx), _ptr(dgamma), _ptr(dbeta)), "stk_layernorm_bwd")
        return dx
    masked = drop is not None and drop.thr > 0
    dxm = torch.empty_like(x) if masked else None
    check(_lib.load().stk_layernorm_bwd_fused(dev, stream, _ptr(dy), _ptr(x), M, _ptr(gamma), _ptr(mean), _ptr(rstd),
                                              _ptr(dx), _ptr(dgamma), _ptr(dbeta), _ptr(dxm), _ptr(dbias),
                                              drop.seed if masked else 0, drop.site if masked else 0,
                                              drop.thr if masked else 0), "stk_layernorm_bwd_fused")
    return (dx, dxm if masked else dx) if drop is not None else dx


# --------------------------------------------------------------------------------------------------
# GEMM
# --------------------------------------------------------------------------------------------------
def gemm(a: torch.Tensor, b: torch.Tensor, *, M: int, N: int, K: int, a_major: int = 0, b_major: int = 0,
         epilogue: int = EPI_BIAS, out: Optional[torch.Tensor] = None, bias=None, resid=None, c2=None, labels=None,
         lse=None, scale_dev=None, ce_partial=None, tgt_logit=None, n_offset: int = 0, split_k: int = 1,
         ln_gamma=None, ln_beta=None, ln_mean=None, ln_rstd=None, drop: Optional["Drop"] = None):
    """C[M,N] = epilogue(A[M,K] B[N,K]^T); a/b are 2-D bf16 tensors in the stored layout
    (K-major: [M,K] / [N,K]; MN-major: [K,M] / [K,N]); row stride = leading dimension."""
    _req(a, torch.bfloat16, "A")
    _req(b, torch.bfloat16, "B")
    assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
    dev, stream = _ctx(a)
    epi = GemmEpilogue()
    epi.bias = bias.data_ptr() if bias is not None else None
    if resid is not None:
        assert resid.stride(1) == 1
        epi.resid = resid.data_ptr()
        epi.ldr = resid.stride(0)
    if c2 is not None:
        epi.c2 = c2.data_ptr()
        epi.ldc2 = c2.stride(0)
    if labels is not None:
        _req(labels, torch.int32, "labels")
        epi.labels = labels.data_ptr()
    epi.lse = lse.data_ptr() if lse is not None else None
    epi.scale_dev = scale_dev.data_ptr() if scale_dev is not None else None
    if ce_partial is not None:
        epi.ce_partial = ce_partial.data_ptr()
        epi.ce_pitch = ce_partial.shape[1]
    epi.tgt_logit = tgt_logit.data_ptr() if tgt_logit is not None else None
    epi.n_offset = n_offset
    epi.ln_gamma = ln_gamma.data_ptr() if ln_gamma is not None else None
    epi.ln_beta = ln_beta.data_ptr() if ln_beta is not None else None
    epi.ln_mean = ln_mean.data_ptr() if ln_mean is not None else None
    epi.ln_rstd = ln_rstd.data_ptr() if ln_rstd is not None else None
    if drop is not None:
        epi.drop_seed, epi.drop_site, epi.drop_thr = drop.seed, drop.site, drop.thr
    if epilogue != EPI_CE_STATS:
        if out is None:
            dt = torch.float32 if epilogue in _F32_EPIS else torch.bfloat16
            out = torch.empty((M, N), dtype=dt, device=a.device)
        assert out.stride(1) == 1
        ldc = out.stride(0)
    else:
        ldc = 0
    prof = _PROFILER
    if prof is not None:
        e0, e1 = prof.span(f"gemm_a{a_major}b{b_major}_epi{epilogue}_n{N}_k{K}", 2.0 * M * N * K)
        e0.record()
    check(_lib.load().stk_gemm(dev, stream, a_major, b_major, _ptr(a), a.stride(0), _ptr(b), b.stride(0), M, N, K,
                               epilogue, _ptr(out), ldc, ctypes.byref(epi), split_k), "stk_gemm")
    if prof is not None:
        e1.record()
    return out


def linear_resid_ln(x, w, bias, resid, gamma, beta, *, save_for_backward=False, drop: Optional["Drop"] = None):
    """y = LayerNorm(drop(x @ w.T + bias) + resid) * gamma + beta in ONE kernel (HF:294-298, 352-356; ``drop`` = the
    train()-mode dropout of the dense output, masks generated inside the epilogue).
    Returns y, or (y, z, mean, rstd) with z the pre-LayerNorm sum when ``save_for_backward``."""
    M = x.shape[0]
    assert w.shape[0] == H
    z = mean = rstd = None
    if save_for_backward:
        z = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
    dropping = drop is not None and drop.thr > 0
    y = gemm(x, w, M=M, N=H, K=x.shape[1], bias=bias, epilogue=EPI_BIAS_DROP_RESID_LN if dropping else EPI_BIAS_RESID_LN,
             resid=resid, c2=z, ln_gamma=gamma, ln_beta=beta, ln_mean=mean, ln_rstd=rstd, drop=drop if dropping else None)
    return (y, z, mean, rstd) if save_for_backward else y


def linear(x, w, bias=None, epilogue=EPI_BIAS, **kw):
    """y = epilogue(x @ w.T + bias) for x [M,K], w [N,K] (nn.Linear layout)."""
    return gemm(x, w, M=x.shape[0], N=w.shape[0], K=x.shape[1], bias=bias, epilogue=epilogue, **kw)


# --------------------------------------------------------------------------------------------------
# attention
# --------------------------------------------------------------------------------------------------
class Drop:
    """One dropout site of a training step: (seed, site id, threshold = round(128 p)); see csrc/stk_rng.cuh."""
    __slots__ = ("seed", "site", "thr")

    def __init__(self, seed: int, site: int, p: float):
        self.seed = seed & 0xFFFFFFFF
        self.site = site & 0xFFFFFFFF
        self.thr = min(127, int(round(128.0 * p)))


def dropout(x: torch.Tensor, d: Drop, out=None) -> torch.Tensor:
    """y = drop(x) over bf16 [M, 768]; the same call is the backward of the site."""
    _req(x, torch.bfloat16, "x")
    assert x.shape[1] == H and x.is_contiguous()
    dev, stream = _ctx(x)
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().stk_dropout_fwd(dev, stream, _ptr(x), x.shape[0], d.seed, d.site, d.thr, _ptr(y)), "stk_dropout_fwd")
    return y


def dropout_resid_ln(x, resid, gamma, beta, d: Drop, save_for_backward=False):
    """y = LN(drop(x) + resid); returns y or (y, z, mean, rstd)."""
    _req(x, torch.bfloat16, "x")
    M = x.shape[0]
    assert x.shape[1] == H and x.is_contiguous() and resid.is_contiguous()
    dev, stream = _ctx(x)
    y = torch.empty_like(x)
    z = mean = rstd = None
    if save_for_backward:
        z = torch.empty_like(x)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
    check(_lib.load().stk_dropout_resid_ln_fwd(dev, stream, _ptr(x), _ptr(resid), M, _ptr(gamma), _ptr(beta), d.seed, d.site,
                                               d.thr, _ptr(z), _ptr(y), _ptr(mean), _ptr(rstd)), "stk_dropout_resid_ln_fwd")
    return (y, z, mean, rstd) if save_for_backward else y


def attention(qkv: torch.Tensor, key_bias: Optional[torch.Tensor], B: int, S: int, save_lse=False, out=None,
              drop: Optional["Drop"] = None, q_rows: Optional[int] = None):
    """``q_rows`` (multiple of 128): only the first q_rows query rows of every sequence are computed; the other rows of
    the returned context are uninitialised (eval only; see engine.encoder_layer_fwd ``first_row_only``)."""
    _req(qkv, torch.bfloat16, "qkv")
    assert qkv.shape == (B * S, 3 * H) and qkv.is_contiguous()
    dev, stream = _ctx(qkv)
    ctx = torch.empty((B * S, H), dtype=torch.bfloat16, device=qkv.device) if out is None else out
    lse = torch.empty((B, HEADS, S), dtype=torch.float32, device=qkv.device) if save_lse else None
    prof = _PROFILER
    if q_rows is not None and q_rows < S:
        assert drop is None and not save_lse
        if prof is not None:
            e0, e1 = prof.span("attn_fwd_qrows", 4.0 * B * HEADS * q_rows * S * 64)
            e0.record()
        check(_lib.load().stk_attn_fwd_qrows(dev, stream, _ptr(qkv), _ptr(key_bias), B, S, q_rows, _ptr(ctx), None),
              "stk_attn_fwd_qrows")
        if prof is not None:
            e1.record()
        return ctx
    if prof is not None:
        e0, e1 = prof.span("attn_fwd", 4.0 * B * HEADS * S * S * 64)
        e0.record()
    if drop is not None and drop.thr > 0:
        check(_lib.load().stk_attn_fwd_dropout(dev, stream, _ptr(qkv), _ptr(key_bias), B, S, _ptr(ctx), _ptr(lse), drop.seed,
                                               drop.site, drop.thr), "stk_attn_fwd_dropout")
    else:
        check(_lib.load().stk_attn_fwd(dev, stream, _ptr(qkv), _ptr(key_bias), B, S, _ptr(ctx), _ptr(lse)),
              "stk_attn_fwd")
    if prof is not None:
        e1.record()
    return (ctx, lse) if save_lse else ctx


def attention_bwd(qkv, key_bias, B, S, out, dout, lse, drop: Optional["Drop"] = None):
    dev, stream = _ctx(qkv)
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(workspace_bytes(WS_ATTN_BWD, B, S) // 4, dtype=torch.float32, device=qkv.device)
    prof = _PROFILER
    if prof is not None:
        e0, e1 = prof.span("attn_bwd", 10.0 * B * HEADS * S * S * 64)
        e0.record()
    if drop is not None and drop.thr > 0:
        check(_lib.load().stk_attn_bwd_dropout(dev, stream, _ptr(qkv), _ptr(key_bias), B, S, _ptr(out), _ptr(dout), _ptr(lse),
                                               _ptr(ws), _ptr(dqkv), drop.seed, drop.site, drop.thr), "stk_attn_bwd_dropout")
    else:
        check(_lib.load().stk_attn_bwd(dev, stream, _ptr(qkv), _ptr(key_bias), B, S, _ptr(out), _ptr(dout), _ptr(lse),
                                       _ptr(ws), _ptr(dqkv)), "stk_attn_bwd")
    if prof is not None:
        e1.record()
    return dqkv


# --------------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------------
def mask_to_bias(mask: torch.Tensor) -> torch.Tensor:
    _req(mask, torch.int64, "attention_mask")
    mask = mask.contiguous()
    dev, stream = _ctx(mask)
    bias = torch.empty(mask.shape, dtype=torch.float32, device=mask.device)
    check(_lib.load().stk_mask_to_bias(dev, stream, _ptr(mask), mask.numel(), _ptr(bias)), "stk_mask_to_bias")
    return bias


def cast_bf16(src: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(src, torch.float32, "src")
    assert src.is_contiguous()
    dev, stream = _ctx(src)
    dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device) if out is None else out
    check(_lib.load().stk_cast_f32_to_bf16(dev, stream, _ptr(src), _ptr(dst), src.numel()), "stk_cast_f32_to_bf16")
    return dst


def gather_rows(src: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    _req(idx, torch.int32, "idx")
    dev, stream = _ctx(src)
    n = idx.numel()
    dst = torch.empty((n, H), dtype=torch.bfloat16, device=src.device)
    if n:
        check(_lib.load().stk_gather_rows(dev, stream, _ptr(src), _ptr(idx), n, _ptr(dst)), "stk_gather_rows")
    return dst


def scatter_add_rows(src: torch.Tensor, idx: torch.Tensor, dst: torch.Tensor) -> None:
    dev, stream = _ctx(src)
    if idx.numel():
        check(_lib.load().stk_scatter_add_rows(dev, stream, _ptr(src), _ptr(idx), idx.numel(), _ptr(dst)),
              "stk_scatter_add_rows")


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = False) -> torch.Tensor:
    dev, stream = _ctx(x)
    M, N = x.shape
    check(_lib.load().stk_colsum(dev, stream, _ptr(x), x.stride(0), M, N, _ptr(out), int(accumulate)), "stk_colsum")
    return out


def ce_finalize(ce_partial, tgt_logit, M):
    dev, stream = _ctx(ce_partial)
    lse = torch.empty(M, dtype=torch.float32, device=ce_partial.device)
    row_loss = torch.empty(M, dtype=torch.float32, device=ce_partial.device)
    check(_lib.load().stk_ce_finalize(dev, stream, _ptr(ce_partial), ce_partial.shape[1], _ptr(tgt_logit), M,
                                      _ptr(lse), _ptr(row_loss)), "stk_ce_finalize")
    return lse, row_loss


def nsp_head(pooled, w, b, labels=None, err_flag=None):
    dev, stream = _ctx(pooled)
    B = pooled.shape[0]
    logits = torch.empty((B, 2), dtype=torch.float32, device=pooled.device)
    row_loss = torch.empty(B, dtype=torch.float32, device=pooled.device) if labels is not None else None
    check(_lib.load().stk_nsp_head_fwd(dev, stream, _ptr(pooled), B, _ptr(w), _ptr(b), _ptr(labels), _ptr(logits),
                                       _ptr(row_loss), _ptr(err_flag)), "stk_nsp_head_fwd")
    return logits, row_loss


def workspace_bytes(op: int, a: int, b: int) -> int:
    n = int(_lib.load().stk_query_workspace(op, a, b))
    if n < 0:
        check(n, "stk_query_workspace")
    return n


def compact_labels(labels: torch.Tensor, row_pitch: int, col_offset: int, vocab: int, capacity: int, err_flag=None):
    """Device-side label selection (``labels != -100``): returns (rows int32 [capacity], labels int32 [capacity],
    count int32 [1]) with padding entries -1 beyond the count; nothing is read back to the host."""
    _req(labels, torch.int64, "labels")
    assert labels.dim() == 2 and labels.is_contiguous()
    dev, stream = _ctx(labels)
    B, width = labels.shape
    rows = torch.empty(capacity, dtype=torch.int32, device=labels.device)
    labs = torch.empty(capacity, dtype=torch.int32, device=labels.device)
    count = torch.empty(1, dtype=torch.int32, device=labels.device)
    check(_lib.load().stk_compact_labels(dev, stream, _ptr(labels), B, width, row_pitch, col_offset, vocab, capacity,
                                         _ptr(rows), _ptr(labs), _ptr(count), _ptr(err_flag)), "stk_compact_labels")
    return rows, labs, count


def linear_ce_fwd(t_rows: torch.Tensor, w: torch.Tensor, labels_i32: torch.Tensor):
    """Fused decoder GEMM + cross-entropy (no logits).  Returns (lse [R], row_loss [R], loss_count [2] = mean loss, count)."""
    _req(t_rows, torch.bfloat16, "t")
    _req(w, torch.bfloat16, "w")
    _req(labels_i32, torch.int32, "labels")
    assert t_rows.is_contiguous() and w.is_contiguous() and t_rows.shape[1] == H and w.shape[1] == H
    dev, stream = _ctx(t_rows)
    R, V = t_rows.shape[0], w.shape[0]
    nbytes = workspace_bytes(WS_LINEAR_CE_FWD, R, V)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=t_rows.device)
    lse = torch.empty(R, dtype=torch.float32, device=t_rows.device)
    row_loss = torch.empty(R, dtype=torch.float32, device=t_rows.device)
    loss_count = torch.empty(2, dtype=torch.float32, device=t_rows.device)
    prof = _PROFILER
    if prof is not None:
        e0, e1 = prof.span(f"linear_ce_fwd_n{V}", 2.0 * R * V * H)
        e0.record()
    check(_lib.load().stk_linear_ce_fwd(dev, stream, _ptr(t_rows), _ptr(w), R, V, _ptr(labels_i32), _ptr(ws), nbytes,
                                        _ptr(lse), _ptr(row_loss), _ptr(loss_count)), "stk_linear_ce_fwd")
    if prof is not None:
        e1.record()
    return lse, row_loss, loss_count


def linear_ce_bwd(t_rows, w, labels_i32, lse, scale_dev, dT, dW) -> None:
    """dT += dlogit W, dW += dlogit^T t with dlogit = (softmax - onehot) * scale recomputed chunk by chunk."""
    dev, stream = _ctx(t_rows)
    R, V = t_rows.shape[0], w.shape[0]
    assert dT.is_contiguous() and dW.is_contiguous() and dT.dtype == torch.float32 and dW.dtype == torch.float32
    nbytes = workspace_bytes(WS_LINEAR_CE_BWD, R, V)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=t_rows.device)
    prof = _PROFILER
    if prof is not None:
        e0, e1 = prof.span(f"linear_ce_bwd_n{V}", 6.0 * R * V * H)
        e0.record()
    check(_lib.load().stk_linear_ce_bwd(dev, stream, _ptr(t_rows), _ptr(w), R, V, _ptr(labels_i32), _ptr(lse),
                                        _ptr(scale_dev), _ptr(ws), nbytes, _ptr(dT), _ptr(dW)), "stk_linear_ce_bwd")
    if prof is not None:
        e1.record()


def masked_mean_pool(seq: torch.Tensor, attention_mask: Optional[torch.Tensor], B: int, shape: "SeqShape") -> torch.Tensor:
    """fp32 [B, 768]: mean of the last hidden state over the attended tokens of each pair."""
    _req(seq, torch.bfloat16, "seq")
    assert seq.shape == (B * shape.seq_pad, H) and seq.is_contiguous()
    if attention_mask is not None:
        _req(attention_mask, torch.int64, "attention_mask")
        assert attention_mask.shape == (B, shape.seq_len) and attention_mask.is_contiguous()
    dev, stream = _ctx(seq)
    out = torch.empty((B, H), dtype=torch.float32, device=seq.device)
    check(_lib.load().stk_masked_mean_pool(dev, stream, _ptr(seq), _ptr(attention_mask), B, shape.seq_len, shape.seq_pad,
                                           _ptr(out)), "stk_masked_mean_pool")
    return out


def scale_heads(x: torch.Tensor, scales: torch.Tensor, out=None) -> torch.Tensor:
    """y[:, h*64:(h+1)*64] = x[:, h*64:(h+1)*64] * scales[h] (head_mask applied to the attention context)."""
    _req(x, torch.bfloat16, "x")
    _req(scales, torch.float32, "scales")
    assert x.shape[1] == H and x.is_contiguous() and scales.numel() == HEADS
    dev, stream = _ctx(x)
    y = torch.empty_like(x) if out is None else out
    check(_lib.load().stk_scale_heads(dev, stream, _ptr(x), x.shape[0], _ptr(scales), _ptr(y)), "stk_scale_heads")
    return y


def gelu_bwd(dy: torch.Tensor, pre: torch.Tensor) -> torch.Tensor:
    dev, stream = _ctx(dy)
    dx = torch.empty_like(dy)
    check(_lib.load().stk_gelu_bwd(dev, stream, _ptr(dy), _ptr(pre), dy.numel(), _ptr(dx)), "stk_gelu_bwd")
    return dx


def nsp_pool_bwd(pooled, logits, labels, scale_dev, w, dw, db) -> torch.Tensor:
    dev, stream = _ctx(pooled)
    B = pooled.shape[0]
    dpre = torch.empty((B, H), dtype=torch.bfloat16, device=pooled.device)
    check(_lib.load().stk_nsp_pool_bwd(dev, stream, _ptr(pooled), _ptr(logits), _ptr(labels), B, _ptr(scale_dev),
                                       _ptr(w), _ptr(dw), _ptr(db), _ptr(dpre)), "stk_nsp_pool_bwd")
    return dpre


def cls_head(pooled, w, b, labels=None, err_flag=None):
    """Sequence-classification head: logits fp32 [B, L] (+ per-sample CE when labels are given)."""
    dev, stream = _ctx(pooled)
    B, L = pooled.shape[0], w.shape[0]
    logits = torch.empty((B, L), dtype=torch.float32, device=pooled.device)
    row_loss = torch.empty(B, dtype=torch.float32, device=pooled.device) if labels is not None else None
    check(_lib.load().stk_cls_head_fwd(dev, stream, _ptr(pooled), B, L, _ptr(w), _ptr(b), _ptr(labels), _ptr(logits),
                                       _ptr(row_loss), _ptr(err_flag)), "stk_cls_head_fwd")
    return logits, row_loss


def cls_pool_bwd(pooled, logits, labels, scale_dev, w, dw, db) -> torch.Tensor:
    dev, stream = _ctx(pooled)
    B, L = logits.shape
    dpre = torch.empty((B, H), dtype=torch.bfloat16, device=pooled.device)
    ws = torch.empty((B, L), dtype=torch.float32, device=pooled.device)
    check(_lib.load().stk_cls_pool_bwd(dev, stream, _ptr(pooled), _ptr(logits), _ptr(labels), B, L, _ptr(scale_dev),
                                       _ptr(w), _ptr(ws), _ptr(dw), _ptr(db), _ptr(dpre)), "stk_cls_pool_bwd")
    return dpre


def unpack_scale(src_bf16: torch.Tensor, dst_f32: torch.Tensor, scale: float) -> None:
    dev, stream = _ctx(src_bf16)
    check(_lib.load().stk_unpack_scale(dev, stream, _ptr(src_bf16), _ptr(dst_f32), src_bf16.numel(), float(scale)),
          "stk_unpack_scale")
