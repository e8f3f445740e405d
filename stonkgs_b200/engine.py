"""Kernel-level orchestration of the STonKGs hot path (forward and backward).

Everything here is a sequence of libstk.so launches on the current CUDA stream; torch tensors are
only the memory the kernels read and write.  The module mirrors, step by step, what the reference
computes in ``STonKGsForPreTraining.forward`` (stonkgs_model.py:149-258) through HF BERT
(modeling_bert.py: BertEmbeddings :72-112, BertLayer :143-207/:287-298/:330-356, BertPooler
:456-468, BertPredictionHeadTransform :471-485, BertPreTrainingHeads :530-533).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import ops

H = 768
I = 3072
HALF = 256
# bias + residual + LayerNorm fused into the Wo / FFN2 GEMM epilogues (STK_EPI_BIAS_RESID_LN); the switch
# exists only so that profiling tools can time the unfused pair (GEMM + LayerNorm kernel) beside it
FUSED_LN = os.environ.get("STK_FUSED_LN", "1") != "0"


# --------------------------------------------------------------------------------------------------
# device-side weight views
# --------------------------------------------------------------------------------------------------
@dataclass
class LayerWeights:
    wqkv: torch.Tensor   # bf16 [2304, 768]  rows = query | key | value  (HF:158-160)
    bqkv: torch.Tensor   # fp32 [2304]
    wo: torch.Tensor     # bf16 [768, 768]    attention.output.dense (HF:290)
    bo: torch.Tensor
    ln1_g: torch.Tensor
    ln1_b: torch.Tensor
    w1: torch.Tensor     # bf16 [3072, 768]   intermediate.dense (HF:333)
    b1: torch.Tensor
    w2: torch.Tensor     # bf16 [768, 3072]   output.dense (HF:348)
    b2: torch.Tensor
    ln2_g: torch.Tensor
    ln2_b: torch.Tensor


@dataclass
class EncoderWeights:
    word: Optional[torch.Tensor]  # fp32 [V, 768] (LM backbone only; the joint encoder's table is dead weight)
    pos: torch.Tensor             # fp32 [512, 768]
    type_emb: torch.Tensor        # fp32 [2, 768]
    emb_g: torch.Tensor
    emb_b: torch.Tensor
    layers: List[LayerWeights] = field(default_factory=list)
    wp: Optional[torch.Tensor] = None   # bf16 [768, 768] pooler.dense
    bp: Optional[torch.Tensor] = None


@dataclass
class HeadWeights:
    wt: torch.Tensor      # bf16 [768, 768]   cls.predictions.transform.dense
    bt: torch.Tensor
    ln_g: torch.Tensor
    ln_b: torch.Tensor
    w_text: torch.Tensor  # bf16 [V, 768]     cls.predictions.text_decoder (no bias, stonkgs_model.py:47)
    w_ent: torch.Tensor   # bf16 [N, 768]     cls.predictions.entity_decoder (no bias, :49)
    w_nsp: torch.Tensor   # fp32 [2, 768]     cls.seq_relationship
    b_nsp: torch.Tensor


@dataclass
class LayerCache:
    """Activations one encoder layer keeps for its backward pass."""
    x_in: torch.Tensor
    qkv: torch.Tensor
    lse: torch.Tensor
    ctx: torch.Tensor
    z1: torch.Tensor
    mean1: torch.Tensor
    rstd1: torch.Tensor
    x1: torch.Tensor
    u: torch.Tensor
    h: torch.Tensor
    z2: torch.Tensor
    mean2: torch.Tensor
    rstd2: torch.Tensor
    ctx_used: Optional[torch.Tensor] = None   # ctx through head_mask (the Wo GEMM's operand); ctx itself when there is none


# --------------------------------------------------------------------------------------------------
# training-mode dropout: site ids shared with oracle/dropout_oracle.py
# --------------------------------------------------------------------------------------------------
@dataclass
class DropCtx:
    """Dropout state of one forward pass in train() mode: every site derives its mask from (seed, site id) through
    the counter-based function of csrc/stk_rng.cuh, so the backward pass regenerates it instead of storing it.
    Encoder 0 = frozen LM backbone, 1 = joint encoder (HF:110 embeddings, :132 attention probabilities, :297 / :355
    dense outputs before the residual sums)."""
    seed: int
    p_hidden: float
    p_attn: float

    def embeddings(self, enc: int) -> ops.Drop:
        return ops.Drop(self.seed, enc * 64 + 63, self.p_hidden)

    def attention(self, enc: int, layer: int) -> ops.Drop:
        return ops.Drop(self.seed, enc * 64 + layer * 4, self.p_attn)

    def attn_out(self, enc: int, layer: int) -> ops.Drop:
        return ops.Drop(self.seed, enc * 64 + layer * 4 + 1, self.p_hidden)

    def ffn_out(self, enc: int, layer: int) -> ops.Drop:
        return ops.Drop(self.seed, enc * 64 + layer * 4 + 2, self.p_hidden)


# --------------------------------------------------------------------------------------------------
# forward
# --------------------------------------------------------------------------------------------------
def encoder_layer_fwd(x, lw: LayerWeights, B: int, S: int, key_bias, cache: Optional[list] = None,
                      drop: Optional[DropCtx] = None, enc: int = 0, li: int = 0, head_scale=None,
                      first_row_only: bool = False):
    """One BertLayer (post-LN). x: bf16 [B*S, 768].

    eval() and train() take the same kernels: in train() the attention kernel drops probabilities (HF:132) and the
    dense outputs are dropped inside the fused bias + residual + LayerNorm GEMM epilogue (HF:296-298, 354-356), masks
    being a pure function of (seed, site, row, column).  ``head_scale`` (fp32 [12], optional) is this layer's row of the
    reference's ``head_mask``: a per-head factor on the attention probabilities, i.e. on the context columns.

    ``first_row_only`` (eval, last layer of the extraction path): the caller only reads row 0 of every sequence
    (BertPooler, HF:456-468), so everything after the K / V projection runs on those B rows — same kernels, same
    per-row arithmetic: the attention computes the first 128-query tile of every (head, pair), the three GEMMs read
    row b*S through the operand pitch.  Returns bf16 [B, 768]."""
    M = x.shape[0]
    train = cache is not None
    qkv = ops.linear(x, lw.wqkv, lw.bqkv)
    if first_row_only:
        assert not train and drop is None and head_scale is None and FUSED_LN
        ctx = ops.attention(qkv, key_bias, B, S, q_rows=128)
        x0 = x.view(B, S, H)[:, 0]
        x1 = ops.linear_resid_ln(ctx.view(B, S, H)[:, 0], lw.wo, lw.bo, x0, lw.ln1_g, lw.ln1_b)
        h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU)
        return ops.linear_resid_ln(h, lw.w2, lw.b2, x1, lw.ln2_g, lw.ln2_b)
    d_attn = drop.attention(enc, li) if drop is not None else None
    d_o = drop.attn_out(enc, li) if drop is not None else None
    d_f = drop.ffn_out(enc, li) if drop is not None else None
    if train:
        ctx, lse = ops.attention(qkv, key_bias, B, S, save_lse=True, drop=d_attn)
    else:
        ctx, lse = ops.attention(qkv, key_bias, B, S, drop=d_attn), None
    ctx_used = ctx if head_scale is None else ops.scale_heads(ctx, head_scale)
    if FUSED_LN:
        if train:
            x1, z1, mean1, rstd1 = ops.linear_resid_ln(ctx_used, lw.wo, lw.bo, x, lw.ln1_g, lw.ln1_b, save_for_backward=True,
                                                       drop=d_o)
            u = torch.empty((M, I), dtype=torch.bfloat16, device=x.device)
            h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU_SAVE_GRAD, c2=u)   # u = gelu'(pre-activation)
            x2, z2, mean2, rstd2 = ops.linear_resid_ln(h, lw.w2, lw.b2, x1, lw.ln2_g, lw.ln2_b, save_for_backward=True,
                                                       drop=d_f)
            cache.append(LayerCache(x, qkv, lse, ctx, z1, mean1, rstd1, x1, u, h, z2, mean2, rstd2, ctx_used))
        else:
            x1 = ops.linear_resid_ln(ctx_used, lw.wo, lw.bo, x, lw.ln1_g, lw.ln1_b, drop=d_o)
            h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU)
            x2 = ops.linear_resid_ln(h, lw.w2, lw.b2, x1, lw.ln2_g, lw.ln2_b, drop=d_f)
        return x2
    # STK_FUSED_LN=0 (profiling tools only): dense GEMM, then a row kernel for (dropout +) residual + LayerNorm
    if drop is not None:
        d1 = ops.linear(ctx_used, lw.wo, lw.bo)
        if train:
            x1, z1, mean1, rstd1 = ops.dropout_resid_ln(d1, x, lw.ln1_g, lw.ln1_b, d_o, save_for_backward=True)
            u = torch.empty((M, I), dtype=torch.bfloat16, device=x.device)
            h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU_SAVE_GRAD, c2=u)
        else:
            x1 = ops.dropout_resid_ln(d1, x, lw.ln1_g, lw.ln1_b, d_o)
            h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU)
        d2 = ops.linear(h, lw.w2, lw.b2)
        if train:
            x2, z2, mean2, rstd2 = ops.dropout_resid_ln(d2, x1, lw.ln2_g, lw.ln2_b, d_f, save_for_backward=True)
            cache.append(LayerCache(x, qkv, lse, ctx, z1, mean1, rstd1, x1, u, h, z2, mean2, rstd2, ctx_used))
        else:
            x2 = ops.dropout_resid_ln(d2, x1, lw.ln2_g, lw.ln2_b, d_f)
        return x2
    z1 = ops.linear(ctx_used, lw.wo, lw.bo, ops.EPI_BIAS_RESID, resid=x)
    if train:
        x1, mean1, rstd1 = ops.layernorm(z1, lw.ln1_g, lw.ln1_b, save_stats=True)
        u = torch.empty((M, I), dtype=torch.bfloat16, device=x.device)
        h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU_SAVE_GRAD, c2=u)   # u = gelu'(pre-activation)
    else:
        x1 = ops.layernorm(z1, lw.ln1_g, lw.ln1_b, out=z1)
        h = ops.linear(x1, lw.w1, lw.b1, ops.EPI_BIAS_GELU)
    z2 = ops.linear(h, lw.w2, lw.b2, ops.EPI_BIAS_RESID, resid=x1)
    if train:
        x2, mean2, rstd2 = ops.layernorm(z2, lw.ln2_g, lw.ln2_b, save_stats=True)
        cache.append(LayerCache(x, qkv, lse, ctx, z1, mean1, rstd1, x1, u, h, z2, mean2, rstd2, ctx_used))
    else:
        x2 = ops.layernorm(z2, lw.ln2_g, lw.ln2_b, out=z2)
    return x2


def encoder_fwd(x, ew: EncoderWeights, B: int, S: int, key_bias, cache: Optional[list] = None,
                drop: Optional[DropCtx] = None, enc: int = 0, head_mask=None, first_row_only: bool = False):
    """``head_mask``: fp32 [layers, 12] on the device, or None.  ``first_row_only``: the LAST layer produces row 0 of
    every sequence only (bf16 [B, 768] instead of [B*S, 768])."""
    last = len(ew.layers) - 1
    for li, lw in enumerate(ew.layers):
        x = encoder_layer_fwd(x, lw, B, S, key_bias, cache, drop, enc, li,
                              head_mask[li] if head_mask is not None else None,
                              first_row_only=first_row_only and li == last)
    return x


def lm_backbone_fwd(ew: EncoderWeights, text_ids: torch.Tensor, key_bias=None, err_flag=None,
                    drop: Optional[DropCtx] = None):
    """``self.lm_backbone(input_ids[:, :256])[0]`` (stonkgs_model.py:178): ids only, no mask, types 0.
    In train() the frozen backbone is in training mode too (the reference never calls ``.eval()`` on it)."""
    B, S = text_ids.shape
    x = ops.embed_text_ln(text_ids, ew.word, ew.pos, ew.type_emb, ew.emb_g, ew.emb_b, err_flag=err_flag)
    if drop is not None:
        ops.dropout(x, drop.embeddings(0), out=x)
    return encoder_fwd(x, ew, B, S, key_bias, None, drop, 0)


def lm_special_rows(ew: EncoderWeights, token_ids) -> torch.Tensor:
    """``lm_backbone(tensor([[id]]))[0][0][0]`` for each id (stonkgs_model.py:138-141).

    A one-token sequence is run as a 128-token padded sequence whose keys 1..127 are masked, which
    is the same arithmetic for row 0 (softmax over one key).  Returns fp32 [len(token_ids), 768]."""
    dev = ew.pos.device
    n = len(token_ids)
    ids = torch.zeros((n, 128), dtype=torch.int64, device=dev)
    ids[:, 0] = torch.tensor(list(token_ids), dtype=torch.int64, device=dev)
    mask = torch.zeros((n, 128), dtype=torch.int64, device=dev)
    mask[:, 0] = 1
    out = lm_backbone_fwd(ew, ids, ops.mask_to_bias(mask))
    return out.view(n, 128, H)[:, 0].float()


def plan_live_rows(mask: np.ndarray, tile: int = 128):
    """Host-side plan of the ``skip_padding`` extraction pass.  ``mask``: the attention mask [B, S] (0 = padding).

    A padded row is never attended to as a key (additive bias finfo.min: its probability is exactly 0) and, as a query,
    only feeds its own output row, which the pooler does not read.  Keys may be visited in any order, so every pair's
    rows are reordered — row 0 ([CLS]) first, then the attended rows in their original order, then the padding — and
    the pair is cut after the first multiple of ``tile`` rows that holds all of them (a pair that needs all S rows is
    left in its original order: bit-identical to the full pass).  Pairs are grouped by that length.

    Returns a list of ``(S_b, pair_idx int64 [n_b], row_idx int32 [n_b * S_b])`` in decreasing ``S_b``: ``row_idx`` are the
    source rows (``pair * S + position``) of the group's packed activations.  A pair without a single attended key keeps
    all S rows (the reference then attends uniformly over every key)."""
    mask = np.asarray(mask)
    B, S = mask.shape
    key = (mask != 0).astype(np.int8)
    attended = key.sum(axis=1)
    key[:, 0] = 2
    order = np.argsort(-key, axis=1, kind="stable").astype(np.int32)          # [B, S] positions, live first
    live = (key > 0).sum(axis=1)
    length = np.minimum(((live + tile - 1) // tile) * tile, S)
    length[attended == 0] = S
    if S % tile:
        length[:] = S
    order[length == S] = np.arange(S, dtype=np.int32)                          # nothing to cut: keep the pair as it is
    groups = []
    base = (np.arange(B, dtype=np.int32) * S)[:, None]
    for sb in sorted(set(length.tolist()), reverse=True):
        pairs = np.nonzero(length == sb)[0]
        rows = (base[pairs] + order[pairs, :sb]).reshape(-1)
        groups.append((int(sb), pairs.astype(np.int64), np.ascontiguousarray(rows, dtype=np.int32)))
    return groups


def joint_fwd(bert: EncoderWeights, input_ids, token_type_ids, attention_mask, lm_hidden, kg_table, *,
              cache: Optional[dict] = None, want_inputs_embeds=False, err_flag=None, drop: Optional[DropCtx] = None,
              shape: ops.SeqShape = ops.STONKGS_SHAPE, head_mask=None, pooled_only: bool = False, live_plan=None):
    """KG lookup + concat + joint embeddings + 12 layers + pooler (stonkgs_model.py:182-212).
    Activations hold ``shape.seq_pad`` rows per pair (== the sequence length for STonKGs; the 260-token TransE variant is
    padded to 384 rows whose tail is masked out as attention keys).

    ``pooled_only`` (eval callers that read ``pooler_output`` alone): the last layer runs on the [CLS] rows only and the
    returned ``seq`` is None; the pooled output is the same arithmetic per row as the full pass.

    ``live_plan`` (eval, pooler only; :func:`plan_live_rows` of the host copy of ``attention_mask``): the joint encoder
    runs on each group's packed rows instead of all ``B * S``; ``seq`` is None."""
    B = input_ids.shape[0]
    S, SP = shape.seq_len, shape.seq_pad
    train = cache is not None
    x, mean, rstd, emb = ops.embed_joint_ln(input_ids, token_type_ids, lm_hidden, kg_table, bert.pos, bert.type_emb,
                                            bert.emb_g, bert.emb_b, save_stats=train,
                                            want_inputs_embeds=want_inputs_embeds, err_flag=err_flag, shape=shape)
    if drop is not None:
        ops.dropout(x, drop.embeddings(1), out=x)
    if SP > S:   # the padding rows must never be attended to
        am = torch.zeros((B, SP), dtype=torch.int64, device=input_ids.device)
        if attention_mask is not None:
            am[:, :S] = attention_mask
        else:
            am[:, :S] = 1
        attention_mask = am
    key_bias = ops.mask_to_bias(attention_mask) if attention_mask is not None else None
    layer_cache = [] if train else None
    pooled_only = pooled_only and not train and drop is None and head_mask is None and FUSED_LN
    if live_plan is not None and len(live_plan) == 1 and live_plan[0][0] == SP:
        live_plan = None                                                           # no pair can be cut: the plain pass
    if live_plan is not None and not train and drop is None and head_mask is None and key_bias is not None and SP == S:
        pooled = torch.empty((B, H), dtype=torch.float32, device=x.device)
        flat_bias = key_bias.view(-1)
        for sb, pairs, rows in live_plan:
            nb = int(pairs.shape[0])
            # pinned staging: a copy from pageable memory would make the host wait for the stream
            rows_d = torch.from_numpy(rows).pin_memory().to(x.device, non_blocking=True)
            xb = ops.gather_rows(x, rows_d)                                        # packed activations [nb * sb, 768]
            kb = torch.index_select(flat_bias, 0, rows_d).view(nb, sb)             # the same rows of the key bias
            out = encoder_fwd(xb, bert, nb, sb, kb, None, None, 1, None, first_row_only=pooled_only)
            cls = out if pooled_only else out.view(nb, sb, H)[:, 0]                 # [CLS] is packed row 0 of every pair
            pb = ops.gemm(cls, bert.wp, M=nb, N=H, K=H, epilogue=ops.EPI_BIAS_TANH_F32, bias=bert.bp)
            pooled.index_copy_(0, torch.from_numpy(pairs).pin_memory().to(x.device, non_blocking=True), pb)
        return None, pooled, emb
    seq = encoder_fwd(x, bert, B, SP, key_bias, layer_cache, drop, 1, head_mask, first_row_only=pooled_only)
    # BertPooler: tanh(W h[:, 0] + b); rows b*SP are read in place through the A pitch
    cls = seq if pooled_only else seq.view(B, SP, H)[:, 0]
    pooled = ops.gemm(cls, bert.wp, M=B, N=H, K=H, epilogue=ops.EPI_BIAS_TANH_F32, bias=bert.bp)
    if pooled_only:
        seq = None
    if train:
        cache.update(emb_mean=mean, emb_rstd=rstd, key_bias=key_bias, layers=layer_cache, lm_hidden=lm_hidden, drop=drop,
                     shape=shape, head_mask=head_mask)
    return seq, pooled, emb
