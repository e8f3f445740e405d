"""Drop-in ``STonKGsForSequenceClassification`` and batched inference.

Reference: ``src/stonkgs/models/stonkgs_finetuning.py:237-346`` (the fine-tuning model: the pre-training
model's embedding stage and joint encoder, then ``dropout -> Linear(768 -> num_labels)`` on the pooled
output) and ``src/stonkgs/api/api.py:318-336`` (``infer_iter``: one forward per row, softmax of the logits).

Everything up to the pooled output is the same kernel sequence as the pre-training model
(``STonKGsForPreTraining.encode``); the head is ``stk_cls_head_fwd`` / ``stk_cls_pool_bwd``.  The module tree
(``bert``, ``lm_backbone``, ``cls``, ``dropout``, ``classifier``) and therefore the fine-tuned checkpoint
layout are the reference's.  Deviations, both documented in DESIGN.md: only the single-label
classification loss of the reference's three ``problem_type`` branches is implemented in the CUDA head (it is
the one every fine-tuning task of the reference uses); training mode applies no dropout.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
from transformers import BertModel
from transformers.modeling_outputs import SequenceClassifierOutput

from . import ops
from ._lib import StkError
from .model import STonKGsForPreTraining


class STonKGsForSequenceClassification(STonKGsForPreTraining):
    """Fine-tuning model of the reference (stonkgs_finetuning.py:237-346), B200-native compute."""

    def __init__(self, config=None, **kwargs):
        num_labels = getattr(config, "num_labels", None) if config is not None else None
        num_labels = kwargs.pop("num_labels", num_labels) or 2
        super().__init__(config, **kwargs)
        self.num_labels = int(num_labels)
        self.config.num_labels = self.num_labels
        if not 1 <= self.num_labels <= 32:
            raise StkError(f"the classification head kernel supports 1..32 labels (got {self.num_labels})")
        self.bert = BertModel(self.config)                      # reference :250 (re-created, then init_weights)
        self.dropout = torch.nn.Dropout(self.config.hidden_dropout_prob)
        self.classifier = torch.nn.Linear(self.config.hidden_size, self.num_labels)
        self.init_weights()
        self._dev_state = None
        self._grad_buffer = None

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, position_ids=None, head_mask=None,
                inputs_embeds=None, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        """Same contract as the reference forward (stonkgs_finetuning.py:257-337)."""
        if position_ids is not None or inputs_embeds is not None:
            raise StkError("position_ids / inputs_embeds are not supported by the fused path "
                           "(the reference ignores both as well)")
        self._raise_on_bad_ids()   # a flag left by the previous training step
        if labels is not None:
            if self.config.problem_type is None:
                if self.num_labels > 1 and labels.dtype in (torch.long, torch.int):
                    self.config.problem_type = "single_label_classification"
                elif self.num_labels == 1:
                    self.config.problem_type = "regression"
                else:
                    self.config.problem_type = "multi_label_classification"
            if self.config.problem_type != "single_label_classification":
                raise StkError("only single_label_classification is implemented in the CUDA classification head")
        # any live, trainable parameter ties the Function into the autograd graph (frozen ones are skipped)
        anchor = next((p for p, _ in self.grad_buffer().param_views), None) if labels is not None and torch.is_grad_enabled() else None
        grad = anchor is not None
        if grad:
            loss, logits = _FinetuneStep.apply(self, (input_ids, attention_mask, token_type_ids, labels, head_mask), anchor)
            self._stage_err_flag()     # checked by FusedAdamW.step / the next forward, after backward is enqueued
        else:
            with torch.no_grad():
                _, pooled, _ = self.encode(input_ids, attention_mask, token_type_ids, head_mask=head_mask)
                loss, logits = _head_fwd(self, pooled, labels, None)
            self._raise_on_bad_ids()
        if not return_dict:
            return ((loss, logits) if loss is not None else (logits,))
        return SequenceClassifierOutput(loss=loss, logits=logits, hidden_states=None, attentions=None)

    @torch.no_grad()
    def predict_proba(self, input_ids, attention_mask=None, token_type_ids=None, err_flag=None,
                      cls_rows_only: bool = False) -> torch.Tensor:
        """softmax(logits) for a batch: what ``infer_iter`` of the reference yields row by row.  The classifier reads the
        pooled [CLS] row only, so ``cls_rows_only`` applies as in :meth:`STonKGsForPreTraining.embed` (same bits)."""
        _, pooled, _ = self.encode(input_ids, attention_mask, token_type_ids, err_flag=err_flag,
                                   pooled_only=cls_rows_only and not self.training)
        logits, _ = ops.cls_head(pooled, self.classifier.weight.data, self.classifier.bias.data)
        return torch.softmax(logits, dim=1)


def _head_fwd(model, pooled, labels, cache: Optional[dict]):
    dev = pooled.device
    err = None
    if labels is not None:
        err = getattr(model, "_pending_err", None)
        if err is None:
            err = model._pending_err = torch.zeros(1, dtype=torch.int32, device=dev)
    lab = labels.to(dev, torch.int64, non_blocking=True).contiguous().view(-1) if labels is not None else None
    if lab is not None and not labels.is_cuda and lab.numel():
        lo, hi = int(labels.min()), int(labels.max())
        if lo < 0 or hi >= model.num_labels:
            raise IndexError(f"label outside [0, {model.num_labels})")
    logits, row_loss = ops.cls_head(pooled, model.classifier.weight.data, model.classifier.bias.data, lab, err)
    loss = row_loss.mean() if row_loss is not None else None
    if cache is not None:
        cache.update(pooled=pooled, cls_logits=logits, cls_labels=lab)
    return loss, logits


class _FinetuneStep(torch.autograd.Function):
    """loss = f(live parameters); gradients go straight into the flat gradient buffer (see training.py)."""

    @staticmethod
    def forward(ctx, model, batch, anchor):
        cache: dict = {}
        input_ids, attention_mask, token_type_ids, labels, head_mask = batch
        seq, pooled, _ = model.encode(input_ids, attention_mask, token_type_ids, cache=cache, head_mask=head_mask)
        cache["seq"] = seq
        loss, logits = _head_fwd(model, pooled, labels, cache)
        ctx.model, ctx.cache, ctx.st = model, cache, model._dev_state
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, *unused):
        from . import training
        model = ctx.model
        gb = model.grad_buffer()
        mode = gb.prepare()
        dp = getattr(model, "_dp", None)
        if dp is not None:
            dp.begin(gb)
        training.backward_classifier(model, ctx.st, ctx.cache, dloss, gb, on_ready=dp.on_ready if dp is not None else None)
        if dp is not None:
            dp.finish(gb)
        gb.publish(mode)
        ctx.cache = None
        return None, None, None


def infer_arrays(model: STonKGsForSequenceClassification, input_ids: np.ndarray,
                 attention_mask: Optional[np.ndarray] = None, token_type_ids: Optional[np.ndarray] = None,
                 batch_size: int = 256, cls_rows_only: bool = False) -> np.ndarray:
    """Class probabilities [n, num_labels] for host id arrays [n, 512] (batched ``infer_iter``), streamed through the
    same reusable pinned staging ring as the embedding extraction (embeddings.EmbeddingStreamer)."""
    from .embeddings import EmbeddingStreamer
    fn = model.predict_proba
    if cls_rows_only:
        def fn(*cols, err_flag=None):
            return model.predict_proba(*cols, err_flag=err_flag, cls_rows_only=True)
    st = EmbeddingStreamer(model, batch_size, fn=fn, out_width=model.num_labels)
    return st.run(input_ids, attention_mask, token_type_ids)


def infer_iter(model: STonKGsForSequenceClassification, rows: Iterable[dict],
               batch_size: int = 256) -> Iterable[Tuple[np.ndarray, List[float]]]:
    """``api.infer_iter`` over already pre-processed rows (dicts with ``input_ids`` / ``attention_mask`` /
    ``token_type_ids``, the output of the reference's ``preprocess_df_for_embeddings_iter``): yields
    ``(logits-free probability row, probability list)`` in input order, computed in batches."""
    buf: List[dict] = []

    def flush():
        ids = np.asarray([r["input_ids"] for r in buf], dtype=np.int64)
        mask = np.asarray([r["attention_mask"] for r in buf], dtype=np.int64) if "attention_mask" in buf[0] else None
        types = np.asarray([r["token_type_ids"] for r in buf], dtype=np.int64) if "token_type_ids" in buf[0] else None
        return infer_arrays(model, ids, mask, types, batch_size)

    for r in rows:
        buf.append(r)
        if len(buf) == batch_size:
            for p in flush():
                yield p, p.tolist()
            buf = []
    if buf:
        for p in flush():
            yield p, p.tolist()


def classes_from_proba(proba: np.ndarray, class_labels: Optional[Sequence[str]] = None):
    """argmax helper mirroring how api.py:338-360 turns probabilities into predicted classes."""
    idx = proba.argmax(axis=1)
    return [class_labels[i] for i in idx] if class_labels is not None else idx
