"""Heads, losses and the hand-written backward pass of ``STonKGsForPreTraining``.

Forward restates stonkgs_model.py:212-258 (ELM head on labelled rows only, three mean
cross-entropies); backward is the autograd backward of the whole trainable path — heads, pooler/NSP,
12 joint encoder layers, joint embedding stage — as an explicit sequence of libstk.so launches.
Gradients are written into one flat fp32 buffer laid out in *reverse execution order* (entity decoder
first, embeddings last), of which every ``param.grad`` is a view; the data-parallel layer all-reduces
slices of that buffer while the rest of backward is still running (``dp.py``).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import engine, ops
from ._lib import StkError

H = 768
I = 3072
HALF = 256
IGNORE = -100


# --------------------------------------------------------------------------------------------------
# flat gradient buffer
# --------------------------------------------------------------------------------------------------
class GradBuffer:
    """One flat fp32 buffer holding the gradients of all live parameters.

    Order = order in which backward produces them, so that bucket k is complete (and can be
    all-reduced) long before backward ends.  Segments are padded to 8 elements: 32 B here (TMA needs 16) and 16 B in
    the bf16 wire buffer of the data-parallel path, whose bucket slices start at segment boundaries."""

    PAD = 8

    def __init__(self, model):
        self.model = model
        dev = model.bert.pooler.dense.weight.device
        pr = model.cls.predictions
        bert = model.bert
        entries = []  # (name, shape, [params that view into it])

        def add(name, params, shape=None):
            shape = tuple(params[0].shape) if shape is None else shape
            entries.append((name, shape, params))

        if getattr(model, "classifier", None) is not None:
            # fine-tuning model (stonkgs_finetuning.py:237-346): only the classifier sits on top of the pooler;
            # the pre-training heads stay in the checkpoint but receive no gradient
            add("cls_w", [model.classifier.weight])
            add("cls_b", [model.classifier.bias])
        else:
            add("w_ent", [pr.entity_decoder.weight])
            add("w_text", [pr.text_decoder.weight])
            add("t_w", [pr.transform.dense.weight])
            add("t_b", [pr.transform.dense.bias])
            add("t_ln_g", [pr.transform.LayerNorm.weight])
            add("t_ln_b", [pr.transform.LayerNorm.bias])
            add("nsp_w", [model.cls.seq_relationship.weight])
            add("nsp_b", [model.cls.seq_relationship.bias])
        add("pool_w", [bert.pooler.dense.weight])
        add("pool_b", [bert.pooler.dense.bias])
        self.num_layers = len(bert.encoder.layer)
        for li in reversed(range(self.num_layers)):
            l = bert.encoder.layer[li]
            a = l.attention.self
            add(f"l{li}.ln2_g", [l.output.LayerNorm.weight])
            add(f"l{li}.ln2_b", [l.output.LayerNorm.bias])
            add(f"l{li}.b2", [l.output.dense.bias])
            add(f"l{li}.w2", [l.output.dense.weight])
            add(f"l{li}.b1", [l.intermediate.dense.bias])
            add(f"l{li}.w1", [l.intermediate.dense.weight])
            add(f"l{li}.ln1_g", [l.attention.output.LayerNorm.weight])
            add(f"l{li}.ln1_b", [l.attention.output.LayerNorm.bias])
            add(f"l{li}.bo", [l.attention.output.dense.bias])
            add(f"l{li}.wo", [l.attention.output.dense.weight])
            add(f"l{li}.bqkv", [a.query.bias, a.key.bias, a.value.bias], (3 * H,))
            add(f"l{li}.wqkv", [a.query.weight, a.key.weight, a.value.weight], (3 * H, H))
        e = bert.embeddings
        add("emb_pos", [e.position_embeddings.weight])
        add("emb_type", [e.token_type_embeddings.weight])
        add("emb_g", [e.LayerNorm.weight])
        add("emb_b", [e.LayerNorm.bias])

        total = 0
        self.offsets = {}
        for name, shape, _ in entries:
            n = 1
            for s in shape:
                n *= s
            self.offsets[name] = (total, n, shape)
            total += (n + self.PAD - 1) // self.PAD * self.PAD
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.views: Dict[str, torch.Tensor] = {}
        # (param, view) of the TRAINABLE live parameters only: a frozen tensor (requires_grad=False, e.g. frozen lower
        # layers while fine-tuning) keeps its slot in the buffer — backward writes there unconditionally — but never
        # gets a .grad, is not updated by FusedAdamW and does not count in the clipping norm (torch semantics)
        self.param_views = []
        for name, shape, params in entries:
            off, n, _ = self.offsets[name]
            v = self.flat[off:off + n].view(shape)
            self.views[name] = v
            if len(params) == 1:
                pv = [(params[0], v)]
            else:  # fused q|k|v: consecutive row blocks
                rows = shape[0] // len(params)
                pv = [(p, v[j * rows:(j + 1) * rows]) for j, p in enumerate(params)]
            self.param_views += [(p, w) for p, w in pv if p.requires_grad]
        self.entries = [e[0] for e in entries]
        self._trainable_key = tuple(p.requires_grad for _, _, ps in entries for p in ps)
        self._all_params = [p for _, _, ps in entries for p in ps]

    def stale(self) -> bool:
        """True when requires_grad of a live parameter changed since the buffer was laid out."""
        return tuple(p.requires_grad for p in self._all_params) != self._trainable_key

    def trainable_runs(self):
        """Maximal contiguous [start, end) element ranges of the flat buffer covered by trainable views (one range — the
        whole buffer — when nothing is frozen): what the global-norm kernel sums over."""
        base = self.flat.data_ptr()
        spans = sorted(((v.data_ptr() - base) // 4, v.numel()) for _, v in self.param_views)
        runs = []
        for off, n in spans:
            end = off + (n + self.PAD - 1) // self.PAD * self.PAD   # segment padding is zero: harmless in a sum of squares
            if runs and off <= runs[-1][1]:
                runs[-1][1] = max(runs[-1][1], end)
            else:
                runs.append([off, end])
        return [(a, min(b, self.flat.numel())) for a, b in runs]

    def __getitem__(self, name):
        return self.views[name]

    def prepare(self):
        """Decide between overwrite (all grads None -> zero the buffer) and accumulate (all grads are
        already our views).  Anything else is handled by a scratch pass + add."""
        states = [(p.grad is None, p.grad is not None and p.grad.data_ptr() == v.data_ptr()) for p, v in self.param_views]
        if all(s[0] for s in states):
            self.flat.zero_()
            return "fresh"
        if all(s[1] for s in states):
            return "accumulate"
        # mixed: keep the user's tensors, compute into a zeroed buffer and add afterwards
        self._foreign = [(p, p.grad) for p, _ in self.param_views]
        self.flat.zero_()
        return "foreign"

    def publish(self, mode):
        if mode == "foreign":
            for (p, v), (_, old) in zip(self.param_views, self._foreign):
                if old is None:
                    p.grad = v.clone()
                else:
                    old.add_(v)
                    p.grad = old
            self._foreign = None
            return
        for p, v in self.param_views:
            p.grad = v


# --------------------------------------------------------------------------------------------------
# heads: forward
# --------------------------------------------------------------------------------------------------
def label_capacity(model, labels: torch.Tensor, width: int) -> int:
    """Rows the head kernels are sized for.  Labels still on the host are counted there (no device sync involved);
    labels already on the device are NOT read back: the capacity is ``model.label_capacity`` per pair (default: the
    ``int(0.15 * width)`` = 38 positions per half that the reference's pre-processing labels,
    indra_for_pretraining.py:55-58; every position for the 4-token TransE entity part) times the batch size, and the
    compaction kernel flags a batch that holds more (StkError at the deferred check)."""
    B = labels.shape[0]
    if not labels.is_cuda:
        return int((labels != IGNORE).sum())
    per_pair = getattr(model, "label_capacity", None)
    if per_pair is None:
        per_pair = int(width * 0.15) if width >= 64 else width
    return max(1, min(int(per_pair), width) * B)


def _compact(model, labels: torch.Tensor, col_offset: int, width: int, row_pitch: int, dev, vocab: int, err):
    """(rows int32 [cap], labels int32 [cap], count int32 [1]) of the labelled positions, or None when cap == 0."""
    if labels.dim() != 2 or labels.shape[1] != width:
        raise StkError(f"label tensors must be [B, {width}], got {tuple(labels.shape)}")
    if not labels.is_cuda and labels.numel():
        lab = labels[labels != IGNORE]
        if lab.numel() and (int(lab.min()) < 0 or int(lab.max()) >= vocab):
            raise IndexError(f"label outside [0, {vocab})")
    cap = label_capacity(model, labels, width)
    if cap == 0:
        return None
    labels_d = labels.to(dev, torch.int64, non_blocking=True).contiguous()
    return ops.compact_labels(labels_d, row_pitch, col_offset, vocab, cap, err)


def _ce_forward(t_rows, w, labels_i32):
    """Fused GEMM + cross-entropy statistics over the vocabulary; logits are never materialised."""
    lse, row_loss, _ = ops.linear_ce_fwd(t_rows, w, labels_i32)
    return lse, row_loss


def heads_fwd(model, hw: engine.HeadWeights, seq, pooled, mlm_labels, elm_labels, nsp_labels, cache: Optional[dict]):
    """ELM head on the labelled rows + the three mean cross-entropies (stonkgs_model.py:62-73, 217-245).  No host sync:
    label selection, row counts, means and range checks all stay on the device."""
    dev = seq.device
    V = hw.w_text.shape[0]
    N = hw.w_ent.shape[0]
    sh = model.seq_shape
    err = getattr(model, "_pending_err", None)
    if err is None:
        err = model._pending_err = torch.zeros(1, dtype=torch.int32, device=dev)
    ct = _compact(model, mlm_labels, 0, sh.text_len, sh.seq_pad, dev, V, err)
    ce = _compact(model, elm_labels, sh.text_len, sh.kg_len, sh.seq_pad, dev, N, err)
    Rt = ct[0].numel() if ct is not None else 0
    Re = ce[0].numel() if ce is not None else 0
    R = Rt + Re
    nan = torch.full((), float("nan"), dtype=torch.float32, device=dev)
    mlm_loss = elm_loss = nan
    lse_t = lse_e = lc_t = lc_e = lab_t = lab_e = rows = None
    u = g = t = mean = rstd = hrows = None
    if R:
        rows = torch.cat([c[0] for c in (ct, ce) if c is not None])
        hrows = ops.gather_rows(seq, rows)
        u = torch.empty((R, H), dtype=torch.bfloat16, device=dev)
        g = ops.linear(hrows, hw.wt, hw.bt, ops.EPI_BIAS_GELU_SAVE, c2=u)
        t, mean, rstd = ops.layernorm(g, hw.ln_g, hw.ln_b, save_stats=True)
        if Rt:
            lab_t = ct[1]
            lse_t, _, lc_t = ops.linear_ce_fwd(t[:Rt], hw.w_text, lab_t)
            mlm_loss = lc_t[0]
        if Re:
            lab_e = ce[1]
            lse_e, _, lc_e = ops.linear_ce_fwd(t[Rt:], hw.w_ent, lab_e)
            elm_loss = lc_e[0]
    if not nsp_labels.is_cuda and nsp_labels.numel() and (int(nsp_labels.min()) < 0 or int(nsp_labels.max()) > 1):
        raise IndexError("next_sentence_labels outside {0, 1}")
    nsp_labels_d = nsp_labels.to(dev, torch.int64, non_blocking=True).contiguous()
    nsp_logits, nsp_rl = ops.nsp_head(pooled, hw.w_nsp, hw.b_nsp, nsp_labels_d, err)
    nsp_loss = nsp_rl.mean()
    loss = mlm_loss + elm_loss + nsp_loss
    if cache is not None:
        cache.update(rows=rows, Rt=Rt, Re=Re, lab_t=lab_t, lab_e=lab_e, lse_t=lse_t, lse_e=lse_e, lc_t=lc_t, lc_e=lc_e,
                     hrows=hrows, u=u, g=g, t=t, t_mean=mean, t_rstd=rstd, nsp_logits=nsp_logits, nsp_labels=nsp_labels_d,
                     pooled=pooled, seq=seq)
    return loss, (mlm_loss, elm_loss, nsp_loss), nsp_logits, (lse_t, lse_e)


def dense_prediction_logits(hw: engine.HeadWeights, seq, B, shape: ops.SeqShape = ops.STONKGS_SHAPE):
    """API-parity mode: the full [B,256,V] and [B,256,N] fp32 logits of stonkgs_model.py:62-73."""
    dev = seq.device
    out = []
    for off, width, w in ((0, shape.text_len, hw.w_text), (shape.text_len, shape.kg_len, hw.w_ent)):
        ar = (torch.arange(B, device=dev, dtype=torch.int32)[:, None] * shape.seq_pad +
              torch.arange(width, device=dev, dtype=torch.int32))
        rows = (ar + off).reshape(-1).contiguous()
        h = ops.gather_rows(seq, rows)
        t = ops.layernorm(ops.linear(h, hw.wt, hw.bt, ops.EPI_BIAS_GELU), hw.ln_g, hw.ln_b)
        V = w.shape[0]
        pitch = (V + 3) // 4 * 4
        buf = torch.empty((B * width, pitch), dtype=torch.float32, device=dev)
        ops.gemm(t, w, M=B * width, N=V, K=H, epilogue=ops.EPI_F32, out=buf[:, :V])
        out.append(buf[:, :V].view(B, width, V))
    return tuple(out)


class LazyPredictionLogits:
    """``prediction_logits`` of a training step: the reference's dense pair ``(text [B,256,V], entity [B,256,N])``
    (stonkgs_model.py:73,253), computed the first time it is indexed, iterated, unpacked or detached.  The training loss
    comes from the fused GEMM + cross-entropy and never needs these 80 GFLOP / 1.4 GB per 8 pairs, so a step that does
    not look at them does not pay for them; anything that does (``outputs[1][0]``, ``text, ent = ...``, HF
    ``nested_detach``) gets real tensors.  Materialise before ``optimizer.step()``: afterwards the decoder weights are
    the updated ones (the sequence output is the step's own)."""

    def __init__(self, hw, seq, B, shape):
        self._args = (hw, seq, B, shape)
        self._pair = None

    def materialize(self):
        if self._pair is None:
            with torch.no_grad():
                self._pair = dense_prediction_logits(*self._args)
            self._args = None
        return self._pair

    def __getitem__(self, i):
        return self.materialize()[i]

    def __iter__(self):
        return iter(self.materialize())

    def __len__(self):
        return 2

    def detach(self):
        return tuple(t.detach() for t in self.materialize())

    def __repr__(self):
        return "LazyPredictionLogits(materialized)" if self._pair is not None else "LazyPredictionLogits(pending)"


def dense_logits_bytes(model, B: int) -> int:
    sh = model.seq_shape
    return 4 * B * (sh.text_len * model.config.vocab_size + sh.kg_len * model.config.kg_vocab_size)


# --------------------------------------------------------------------------------------------------
# backward
# --------------------------------------------------------------------------------------------------
def _wgrad(dy, x, out, k_rows):
    """out[Nout, Nin] += dy[k_rows, Nout]^T x[k_rows, Nin]  (both operands read in place, MN-major; split-K chosen by the
    library so that the persistent grid is full)."""
    n_out, n_in = out.shape
    ops.gemm(dy, x, M=n_out, N=n_in, K=k_rows, a_major=1, b_major=1, epilogue=ops.EPI_F32_ADD, out=out, split_k=0)


def _dgrad(dy, w, *, epilogue=ops.EPI_BIAS, resid=None):
    """dx[M, Nin] = dy[M, Nout] w[Nout, Nin]  (w in its nn.Linear layout = MN-major B)."""
    M, n_out = dy.shape
    n_in = w.shape[1]
    return ops.gemm(dy, w, M=M, N=n_in, K=n_out, b_major=1, epilogue=epilogue, resid=resid)


def _ce_backward(t_rows, w, labels_i32, lse, scale_dev, dT, gW):
    """Fused linear + cross-entropy backward (stk_linear_ce_bwd owns the vocabulary-chunk loop)."""
    ops.linear_ce_bwd(t_rows, w, labels_i32, lse, scale_dev, dT, gW)


def backward(model, st, cache, dloss: torch.Tensor, gb: GradBuffer, on_ready=None):
    """Full backward of loss -> all live parameters.  ``on_ready(name)`` is called (host side) right
    after the launches that complete gradient segment ``name`` have been enqueued."""
    bert: engine.EncoderWeights = st["bert"]
    hw: engine.HeadWeights = st["heads"]
    seq, pooled = cache["seq"], cache["pooled"]
    B = pooled.shape[0]
    dev = seq.device
    M = seq.shape[0]
    ready = on_ready or (lambda name: None)
    dloss = dloss.to(dev, torch.float32).reshape(())

    dseq = torch.zeros((M, H), dtype=torch.bfloat16, device=dev)

    # ---- MLM / ELM heads ------------------------------------------------------------------------
    Rt, Re = cache["Rt"], cache["Re"]
    R = Rt + Re
    if R:
        t = cache["t"]
        dT = torch.zeros((R, H), dtype=torch.float32, device=dev)
        # gradient scale = upstream / number of labelled rows (a device scalar: the count is never read back)
        if Re:
            _ce_backward(t[Rt:], hw.w_ent, cache["lab_e"], cache["lse_e"], (dloss / cache["lc_e"][1]).reshape(1), dT[Rt:],
                         gb["w_ent"])
        ready("w_ent")
        if Rt:
            _ce_backward(t[:Rt], hw.w_text, cache["lab_t"], cache["lse_t"], (dloss / cache["lc_t"][1]).reshape(1), dT[:Rt],
                         gb["w_text"])
        ready("w_text")
        dT_bf = ops.cast_bf16(dT)
        dg = ops.layernorm_bwd(dT_bf, cache["g"], hw.ln_g, cache["t_mean"], cache["t_rstd"], gb["t_ln_g"], gb["t_ln_b"])
        du = ops.gelu_bwd(dg, cache["u"])
        ops.colsum(du, gb["t_b"], accumulate=True)
        _wgrad(du, cache["hrows"], gb["t_w"], R)
        dh = _dgrad(du, hw.wt)
        ops.scatter_add_rows(dh, cache["rows"], dseq)
    else:
        ready("w_ent")
        ready("w_text")
    for n in ("t_w", "t_b", "t_ln_g", "t_ln_b"):
        ready(n)

    # ---- NSP head + pooler ----------------------------------------------------------------------
    dpre = ops.nsp_pool_bwd(pooled, cache["nsp_logits"], cache["nsp_labels"], (dloss / B).reshape(1), hw.w_nsp,
                            gb["nsp_w"], gb["nsp_b"])
    for n in ("nsp_w", "nsp_b"):
        ready(n)
    backward_pooler(bert, seq, dpre, dseq, gb, ready)
    backward_trunk(model, bert, cache, dseq, gb, ready)


def backward_pooler(bert: engine.EncoderWeights, seq, dpre, dseq, gb: GradBuffer, ready):
    """Backward of BertPooler's dense layer (HF:456-468) given the gradient w.r.t. its pre-activation;
    the [CLS] rows of ``dseq`` receive the result."""
    B = dpre.shape[0]
    SP = seq.shape[0] // B
    ops.colsum(dpre, gb["pool_b"], accumulate=True)
    seq0 = seq.view(B, SP, H)[:, 0]
    _wgrad(dpre, seq0, gb["pool_w"], B)
    dseq0 = _dgrad(dpre, bert.wp)
    cls_rows = (torch.arange(B, device=seq.device, dtype=torch.int32) * SP).contiguous()
    ops.scatter_add_rows(dseq0, cls_rows, dseq)
    for n in ("pool_w", "pool_b"):
        ready(n)


def backward_trunk(model, bert: engine.EncoderWeights, cache, dseq, gb: GradBuffer, ready):
    """12 joint encoder layers (last to first) and the joint embedding stage, given d(loss)/d(sequence output)."""
    M = dseq.shape[0]
    shape: ops.SeqShape = cache.get("shape", ops.STONKGS_SHAPE)
    SP = shape.seq_pad
    B = M // SP
    key_bias = cache["key_bias"]
    drop: Optional[engine.DropCtx] = cache.get("drop")   # train() with dropout: masks are regenerated from (seed, site)
    head_mask = cache.get("head_mask")
    dx = dseq
    for li in reversed(range(len(bert.layers))):
        lw = bert.layers[li]
        c: engine.LayerCache = cache["layers"][li]
        p = f"l{li}."
        # dz = gradient w.r.t. the residual sum z; the dense branch sees it through the dropout mask, the residual
        # branch (added back below through the dgrad epilogues) sees it as it is
        # one row pass: LayerNorm backward, the dropout mask of the dense output and the dense bias gradient
        if drop is not None:
            dz2, dz2m = ops.layernorm_bwd(dx, c.z2, lw.ln2_g, c.mean2, c.rstd2, gb[p + "ln2_g"], gb[p + "ln2_b"],
                                          dbias=gb[p + "b2"], drop=drop.ffn_out(1, li))
        else:
            dz2 = dz2m = ops.layernorm_bwd(dx, c.z2, lw.ln2_g, c.mean2, c.rstd2, gb[p + "ln2_g"], gb[p + "ln2_b"],
                                           dbias=gb[p + "b2"])
        _wgrad(dz2m, c.h, gb[p + "w2"], M)
        du = _dgrad(dz2m, lw.w2, epilogue=ops.EPI_MUL, resid=c.u)   # c.u holds gelu'(pre-activation), saved by the forward
        ops.colsum(du, gb[p + "b1"], accumulate=True)
        _wgrad(du, c.x1, gb[p + "w1"], M)
        dx1 = _dgrad(du, lw.w1, epilogue=ops.EPI_BIAS_RESID, resid=dz2)
        if drop is not None:
            dz1, dz1m = ops.layernorm_bwd(dx1, c.z1, lw.ln1_g, c.mean1, c.rstd1, gb[p + "ln1_g"], gb[p + "ln1_b"],
                                          dbias=gb[p + "bo"], drop=drop.attn_out(1, li))
        else:
            dz1 = dz1m = ops.layernorm_bwd(dx1, c.z1, lw.ln1_g, c.mean1, c.rstd1, gb[p + "ln1_g"], gb[p + "ln1_b"],
                                           dbias=gb[p + "bo"])
        ctx_used = c.ctx_used if c.ctx_used is not None else c.ctx
        _wgrad(dz1m, ctx_used, gb[p + "wo"], M)
        dctx = _dgrad(dz1m, lw.wo)
        if head_mask is not None:   # ctx_used = ctx * head_mask[layer] per head: the same factor on the way back
            ops.scale_heads(dctx, head_mask[li], out=dctx)
        dqkv = ops.attention_bwd(c.qkv, key_bias, B, SP, c.ctx, dctx, c.lse,
                                 drop=drop.attention(1, li) if drop is not None else None)
        # bias gradients of query and value; the key-bias gradient is analytically zero (softmax is
        # invariant to a per-query shift of the scores), so its segment stays exactly 0
        # (one pass over all 2304 columns, then the key third — which nothing else ever writes — is reset to its exact
        # zero: one column-sum launch per layer less than summing the query and value thirds apart)
        ops.colsum(dqkv, gb[p + "bqkv"], accumulate=True)
        gb[p + "bqkv"][H:2 * H].zero_()
        _wgrad(dqkv, c.x_in, gb[p + "wqkv"], M)
        dx = _dgrad(dqkv, lw.wqkv, epilogue=ops.EPI_BIAS_RESID, resid=dz1)
        for n in ("ln2_g", "ln2_b", "b2", "w2", "b1", "w1", "ln1_g", "ln1_b", "bo", "wo", "bqkv", "wqkv"):
            ready(p + n)

    # ---- joint embedding stage (position / token-type tables and LayerNorm) ----------------------
    if drop is not None:
        dx = ops.dropout(dx, drop.embeddings(1))
    ops.embed_joint_ln_bwd(cache["input_ids"], cache["token_type_ids"], cache["lm_hidden"], model.kg_table, bert.pos,
                           bert.type_emb, bert.emb_g, cache["emb_mean"], cache["emb_rstd"], dx, gb["emb_pos"],
                           gb["emb_type"], gb["emb_g"], gb["emb_b"], shape=shape)
    for n in ("emb_pos", "emb_type", "emb_g", "emb_b"):
        ready(n)


def backward_classifier(model, st, cache, dloss: torch.Tensor, gb: GradBuffer, on_ready=None):
    """Backward of the fine-tuning model: classifier CE -> pooler -> trunk (no pre-training heads)."""
    bert: engine.EncoderWeights = st["bert"]
    seq, pooled = cache["seq"], cache["pooled"]
    B = pooled.shape[0]
    ready = on_ready or (lambda name: None)
    dloss = dloss.to(seq.device, torch.float32).reshape(())
    dseq = torch.zeros_like(seq)
    dpre = ops.cls_pool_bwd(pooled, cache["cls_logits"], cache["cls_labels"], (dloss / B).reshape(1),
                            model.classifier.weight.data, gb["cls_w"], gb["cls_b"])
    for n in ("cls_w", "cls_b"):
        ready(n)
    backward_pooler(bert, seq, dpre, dseq, gb, ready)
    backward_trunk(model, bert, cache, dseq, gb, ready)


# --------------------------------------------------------------------------------------------------
# autograd glue + reference-shaped outputs
# --------------------------------------------------------------------------------------------------
class _PretrainStep(torch.autograd.Function):
    """loss = f(live parameters).  Gradients are written by the hand-written backward directly into
    the flat gradient buffer (``param.grad`` views), so autograd receives ``None`` for them."""

    @staticmethod
    def forward(ctx, model, batch, anchor):
        cache: dict = {}
        input_ids, attention_mask, token_type_ids, mlm, elm, nsp, head_mask = batch
        seq, pooled, _ = model.encode(input_ids, attention_mask, token_type_ids, cache=cache, need_heads=True,
                                      head_mask=head_mask)
        st = model._dev_state
        loss, parts, nsp_logits, _ = heads_fwd(model, st["heads"], seq, pooled, mlm, elm, nsp, cache)
        ctx.model, ctx.cache, ctx.st = model, cache, st
        ctx.mark_non_differentiable(pooled, nsp_logits, seq)
        model._last_loss_parts = parts
        return loss, pooled, nsp_logits, seq

    @staticmethod
    def backward(ctx, dloss, *unused):
        model = ctx.model
        gb = model.grad_buffer()
        mode = gb.prepare()
        dp = getattr(model, "_dp", None)
        if dp is not None:
            dp.begin(gb)
        backward(model, ctx.st, ctx.cache, dloss, gb, on_ready=dp.on_ready if dp is not None else None)
        if dp is not None:
            dp.finish(gb)
        gb.publish(mode)
        ctx.cache = None
        return None, None, None


def forward(model, input_ids, attention_mask, token_type_ids, mlm, elm, nsp, return_dict, head_mask=None):
    """Reference forward contract (stonkgs_model.py:149-258)."""
    from .model import BertForPreTrainingOutputWithPooling
    have_labels = mlm is not None and elm is not None and nsp is not None
    B = input_ids.shape[0]
    model._raise_on_bad_ids()   # a flag left by the previous step (deferred read, see model._raise_on_bad_ids)
    anchor = next((p for p, _ in model.grad_buffer().param_views), None) if have_labels and torch.is_grad_enabled() else None
    grad = anchor is not None   # a live, trainable parameter ties the Function into the autograd graph
    total_loss = None
    if grad:
        total_loss, pooled, nsp_logits, seq = _PretrainStep.apply(
            model, (input_ids, attention_mask, token_type_ids, mlm, elm, nsp, head_mask), anchor)
        hw = model._dev_state["heads"]
        model._stage_err_flag()
    else:
        with torch.no_grad():
            seq, pooled, _ = model.encode(input_ids, attention_mask, token_type_ids, need_heads=True, head_mask=head_mask)
            hw = model._dev_state["heads"]
            if have_labels:
                total_loss, parts, nsp_logits, _ = heads_fwd(model, hw, seq, pooled, mlm, elm, nsp, None)
                model._last_loss_parts = parts
            else:
                nsp_logits, _ = ops.nsp_head(pooled, hw.w_nsp, hw.b_nsp, None)
    if not grad:
        model._raise_on_bad_ids()   # the training step reads the flag after backward has been enqueued
    want = model.return_prediction_logits
    if want is False:
        prediction_scores = (None, None)                      # explicit opt-out
    elif want or (not grad and dense_logits_bytes(model, B) <= model.dense_logits_max_bytes):
        with torch.no_grad():
            prediction_scores = dense_prediction_logits(hw, seq, B, model.seq_shape)
    else:
        prediction_scores = LazyPredictionLogits(hw, seq, B, model.seq_shape)
    sh = model.seq_shape
    sequence_output = seq.view(B, sh.seq_pad, H)
    if sh.seq_pad != sh.seq_len:
        sequence_output = sequence_output[:, :sh.seq_len]
    if return_dict:
        sequence_output = sequence_output.float()  # the reference returns fp32 hidden states
    if not return_dict:
        output = (prediction_scores, nsp_logits)
        return ((total_loss,) + output) if total_loss is not None else output
    return BertForPreTrainingOutputWithPooling(
        loss=total_loss, prediction_logits=prediction_scores, seq_relationship_logits=nsp_logits,
        hidden_states=sequence_output, attentions=None, pooler_output=pooled)
