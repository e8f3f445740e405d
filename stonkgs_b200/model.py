"""Drop-in ``STonKGsForPreTraining`` whose forward/backward run on hand-written sm_100a kernels.

Boundary kept identical to the reference (``src/stonkgs/models/stonkgs_model.py``):

* constructor ``STonKGsForPreTraining(config, nlp_model_type, kg_embedding_dict_path)`` (:79-84),
  ``from_pretrained`` / ``from_default_pretrained`` (:143-147), ``forward(...)`` signature and
  outputs (:149-159, :247-258), ``BertForPreTrainingOutputWithPooling`` (:30-34);
* the module tree — and therefore the 413 ``state_dict`` keys and the HF config (+ ``kg_vocab_size``)
  — is the reference's: ``bert`` (HF BertModel), ``lm_backbone`` (frozen HF BertModel), ``cls`` with
  ``STonKGsELMPredictionHead`` (:37-60).  The HF modules are *parameter containers only*: their
  ``forward`` methods are never called; all arithmetic goes through ``engine`` -> ``libstk.so``;
* attributes other reference code relies on: ``cls.predictions.half_length``, ``lm_backbone``,
  ``kg_backbone`` (mapping id -> vector), ``kg_idx_to_name``, ``lm_sep_id/lm_mask_id/lm_unk_id``.

There is no CPU / PyTorch fallback: calling ``forward`` without CUDA tensors raises.
"""
from __future__ import annotations

import logging
import os
from collections.abc import Mapping
from dataclasses import dataclass
from functools import lru_cache
from typing import Optional

import numpy as np
import torch
from torch import nn
from transformers import BertConfig, BertForPreTraining, BertModel
from transformers.models.bert.modeling_bert import BertForPreTrainingOutput, BertLMPredictionHead

from . import engine, ops
from ._lib import StkError

logger = logging.getLogger(__name__)

NLP_MODEL_TYPE = "dmis-lab/biobert-v1.1"            # reference constants.py:120-124
EMBEDDINGS_PATH = os.environ.get("STONKGS_EMBEDDINGS_PATH", "embeddings_best_model.tsv")
SEP_ID, MASK_ID, UNK_ID = 102, 103, 100            # BioBERT tokenizer ids (stonkgs_model.py:116-118)
H = 768


@dataclass
class BertForPreTrainingOutputWithPooling(BertForPreTrainingOutput):
    """Same extra field as the reference (stonkgs_model.py:30-34)."""

    pooler_output: Optional[torch.FloatTensor] = None


class STonKGsELMPredictionHead(BertLMPredictionHead):
    """Parameter layout of the reference ELM head (stonkgs_model.py:37-60); forward is CUDA-only."""

    def __init__(self, config):
        super().__init__(config)
        self.text_decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.entity_decoder = nn.Linear(config.hidden_size, config.kg_vocab_size, bias=False)
        self.half_length = config.max_position_embeddings // 2
        self.text_bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.entity_bias = nn.Parameter(torch.zeros(config.kg_vocab_size))
        self.decoder.text_bias = self.text_bias
        self.decoder.entity_bias = self.entity_bias

    def forward(self, hidden_states):  # pragma: no cover - guarded
        raise StkError("STonKGsELMPredictionHead.forward is fused into STonKGsForPreTraining.forward")


def prepare_df(path: str):
    """Load the node2vec TSV (``name \\t 768 floats`` per line, reference node2vec.py:350-354) in file
    order.  Restates ``kg_baseline_model.py:270-280`` without the per-row pandas loop.
    Returns (names, float32 [N, 768])."""
    if str(path).endswith(".npy"):      # binary table written by inputs.save_kg_table
        from .inputs import load_kg_table
        return load_kg_table(str(path))
    names, rows = [], []
    with open(path, "r") as f:
        for line in f:
            parts = line.rstrip("\n").split("\t")
            if len(parts) < 2:
                continue
            names.append(parts[0])
            rows.append(np.asarray(parts[1:], dtype=np.float64))
    return names, np.stack(rows).astype(np.float32)


class _KGBackboneView(Mapping):
    """``kg_backbone`` of the reference is a dict id -> vector (stonkgs_model.py:131-141); this is a
    read-only mapping view over the dense table that restates it."""

    def __init__(self, model):
        self._m = model

    def __getitem__(self, i):
        t = self._m.kg_table
        if not (0 <= int(i) < t.shape[0]):
            raise KeyError(i)
        return t[int(i)]

    def __iter__(self):
        return iter(range(self._m.kg_table.shape[0]))

    def __len__(self):
        return self._m.kg_table.shape[0]


class STonKGsForPreTraining(BertForPreTraining):
    """STonKGs pre-training model (text + KG joint transformer), B200-native compute."""

    # The reference registers text_bias / entity_bias a second time on the inherited decoder
    # (stonkgs_model.py:58-60), i.e. the same Parameter under two state-dict keys.  Declaring the
    # aliases lets the installed transformers (safetensors refuses aliased tensors) save and re-tie them.
    _stk_alias_keys = {
        "cls.predictions.decoder.text_bias": "cls.predictions.text_bias",
        "cls.predictions.decoder.entity_bias": "cls.predictions.entity_bias",
    }

    def __init__(self, config=None, nlp_model_type=NLP_MODEL_TYPE, kg_embedding_dict_path=EMBEDDINGS_PATH):
        # --- KG vectors in file order (reference :93) -------------------------------------------
        if isinstance(kg_embedding_dict_path, (str, os.PathLike)):
            names, rows = prepare_df(kg_embedding_dict_path)
        elif isinstance(kg_embedding_dict_path, dict):  # name -> vector, like the reference's dict
            names = list(kg_embedding_dict_path.keys())
            rows = np.stack([np.asarray(v, dtype=np.float32) for v in kg_embedding_dict_path.values()])
        else:  # (names, array) or a bare array: synthetic tables for tests / benchmarks
            if isinstance(kg_embedding_dict_path, tuple):
                names, rows = kg_embedding_dict_path
            else:
                rows = kg_embedding_dict_path
                names = None
            rows = np.ascontiguousarray(rows, dtype=np.float32)
        n_kg = rows.shape[0]

        # --- config (reference :96-97: rebuilt from the LM type; the passed one is only a fallback
        #     when the hub is unreachable, e.g. from_pretrained of a local checkpoint offline) ----
        passed = config
        if isinstance(nlp_model_type, BertConfig):
            lm_config = nlp_model_type                     # offline: random-init LM backbone of this shape
            config = BertConfig.from_dict(lm_config.to_dict())
        else:
            try:
                config = BertConfig.from_pretrained(nlp_model_type)
                lm_config = None
            except Exception:  # noqa: BLE001  (no network / not cached)
                if not isinstance(passed, BertConfig):
                    raise
                config = BertConfig.from_dict(passed.to_dict())
                lm_config = BertConfig.from_dict(passed.to_dict())
        config.update({"kg_vocab_size": n_kg})
        self._adjust_config(config)
        super().__init__(config)
        self.cls.predictions = self._head_class(config)
        # the head swap happens after HF's post_init(): register the two extra aliases now
        self._tied_weights_keys = {**(type(self)._tied_weights_keys or {}), **self._stk_alias_keys}
        if hasattr(self, "get_expanded_tied_weights_keys"):
            self.all_tied_weights_keys = self.get_expanded_tied_weights_keys(all_submodels=False)

        # --- frozen LM backbone (reference :107-114) --------------------------------------------
        if lm_config is None:
            try:
                self.lm_backbone = BertModel.from_pretrained(nlp_model_type)
            except Exception:  # noqa: BLE001
                lm_fallback = BertConfig.from_dict(config.to_dict())   # weights then come from the checkpoint's lm_backbone.*
                lm_fallback.max_position_embeddings = 512              # the LM keeps its own 512 positions in every variant
                self.lm_backbone = BertModel(lm_fallback)
        else:
            self.lm_backbone = BertModel(lm_config)
        for p in self.lm_backbone.parameters():
            p.requires_grad = False
        self.lm_sep_id, self.lm_mask_id, self.lm_unk_id = SEP_ID, MASK_ID, UNK_ID
        if lm_config is None:
            try:  # tokenizer ids, when the tokenizer is available offline (reference :116-118)
                from transformers import BertTokenizer
                tok = BertTokenizer.from_pretrained(nlp_model_type)
                self.lm_sep_id, self.lm_mask_id, self.lm_unk_id = tok.sep_token_id, tok.mask_token_id, tok.unk_token_id
            except Exception:  # noqa: BLE001
                pass
        self._check_shape(config)

        # --- dense KG table restating the reference's index quirk (reference :123-134) ----------
        specials = (self.lm_sep_id, self.lm_mask_id, self.lm_unk_id)
        numeric_indices = [i for i in range(n_kg + 3) if i not in specials][:n_kg]
        if names is None:
            names = [f"n{i}" for i in range(n_kg)]
        self.kg_idx_to_name = dict(zip(numeric_indices, names))
        table = np.zeros((n_kg + 3, H), dtype=np.float32)
        table[np.asarray(numeric_indices, dtype=np.int64)] = rows
        # a plain attribute, not a buffer: the reference does not store it in the checkpoint (fact 5);
        # built through numpy so that it is a real CPU tensor even under HF's meta-device init context
        self.kg_table = torch.from_numpy(table)
        self.kg_backbone = _KGBackboneView(self)
        self._special_rows_version = None
        self._dev_state = None
        # prediction_logits = the reference's dense pair (text [B,256,V], entity [B,256,N]; stonkgs_model.py:73,253).
        # None (default): computed eagerly whenever no autograd step is being recorded (eval / no labels / no_grad) and
        # the pair fits `dense_logits_max_bytes`; otherwise (the training step, whose loss never needs them) a lazy pair
        # that runs the two decoder GEMMs the first time it is indexed / iterated.  True: always eager.  False: explicit
        # opt-out, the field is (None, None).
        self.return_prediction_logits = None
        self.dense_logits_max_bytes = 16 << 30
        # train() mode dropout (HF:110,132,297,355): masks are a counter-based function of
        # (stk_dropout_seed + step, site), so backward regenerates them; last_dropout_seed is what a test hands to
        # oracle.dropout_oracle.DropSpec to reproduce the same step in fp32
        self.label_capacity = None    # labelled positions per pair and half the heads are sized for (None: 15 %)
        self._pending_err = None
        self._err_staged = None
        self._err_stream = None
        self.stk_dropout = True
        self.stk_dropout_seed = int(torch.initial_seed()) & 0xFFFFFFFF
        self._dropout_step = 0
        self.last_dropout_seed = None

    # ------------------------------------------------------------------------------------------
    _head_class = STonKGsELMPredictionHead

    @staticmethod
    def _adjust_config(config):
        """Variant hook (the TransE model changes max_position_embeddings here, transestonkgs_model.py:93)."""

    #: joint sequence of one pair (text tokens, text + KG tokens, activation rows per pair); the TransE variant overrides it
    seq_shape = ops.STONKGS_SHAPE

    @classmethod
    def _check_shape(cls, config):
        if (config.hidden_size, config.num_attention_heads, config.intermediate_size) != (768, 12, 3072) or \
                config.max_position_embeddings != cls.seq_shape.seq_len or config.hidden_act != "gelu":
            raise StkError("stonkgs_b200 kernels are specialised for the BERT-base shape of the reference "
                           f"(hidden 768, 12 heads, intermediate 3072, {cls.seq_shape.seq_len} positions, erf-GELU)")

    @classmethod
    @lru_cache(maxsize=32)
    def from_default_pretrained(cls, **kwargs) -> "STonKGsForPreTraining":
        """Reference stonkgs_model.py:143-147."""
        return cls.from_pretrained("stonkgs/stonkgs-150k", **kwargs)

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        if "kg_table" in self.__dict__:
            self.kg_table = fn(self.kg_table)   # follows .to()/.cuda() like the reference's dict tensors
            self._dev_state = None
            self._special_rows_version = None
            self._grad_buffer = None
        return out

    # ------------------------------------------------------------------------------------------
    # device-side state: bf16 weight copies and the three LM-backbone rows of the KG table
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def _enc_params(bert: BertModel):
        ps = [bert.embeddings.position_embeddings.weight, bert.embeddings.token_type_embeddings.weight,
              bert.embeddings.LayerNorm.weight, bert.embeddings.LayerNorm.bias]
        for l in bert.encoder.layer:
            ps += [l.attention.self.query.weight, l.attention.self.key.weight, l.attention.self.value.weight,
                   l.attention.self.query.bias, l.attention.self.key.bias, l.attention.self.value.bias,
                   l.attention.output.dense.weight, l.attention.output.dense.bias,
                   l.attention.output.LayerNorm.weight, l.attention.output.LayerNorm.bias,
                   l.intermediate.dense.weight, l.intermediate.dense.bias, l.output.dense.weight, l.output.dense.bias,
                   l.output.LayerNorm.weight, l.output.LayerNorm.bias]
        ps += [bert.pooler.dense.weight, bert.pooler.dense.bias]
        return ps

    def _build_encoder_weights(self, bert: BertModel, with_word: bool, old: Optional[engine.EncoderWeights]):
        """bf16 copies of the GEMM weights (cast kernel), fp32 views of everything else."""
        dev = bert.embeddings.position_embeddings.weight.device
        emb = bert.embeddings
        ew = engine.EncoderWeights(
            word=emb.word_embeddings.weight.data if with_word else None,
            pos=emb.position_embeddings.weight.data, type_emb=emb.token_type_embeddings.weight.data,
            emb_g=emb.LayerNorm.weight.data, emb_b=emb.LayerNorm.bias.data)
        for i, l in enumerate(bert.encoder.layer):
            prev = old.layers[i] if old is not None else None
            att = l.attention.self
            wqkv = prev.wqkv if prev else torch.empty((3 * H, H), dtype=torch.bfloat16, device=dev)
            bqkv = prev.bqkv if prev else torch.empty(3 * H, dtype=torch.float32, device=dev)
            for j, lin in enumerate((att.query, att.key, att.value)):
                ops.cast_bf16(lin.weight.data, out=wqkv[j * H:(j + 1) * H])
                bqkv[j * H:(j + 1) * H].copy_(lin.bias.data)

            def c(w, old_t):
                return ops.cast_bf16(w.data, out=old_t)

            ew.layers.append(engine.LayerWeights(
                wqkv=wqkv, bqkv=bqkv,
                wo=c(l.attention.output.dense.weight, prev.wo if prev else None), bo=l.attention.output.dense.bias.data,
                ln1_g=l.attention.output.LayerNorm.weight.data, ln1_b=l.attention.output.LayerNorm.bias.data,
                w1=c(l.intermediate.dense.weight, prev.w1 if prev else None), b1=l.intermediate.dense.bias.data,
                w2=c(l.output.dense.weight, prev.w2 if prev else None), b2=l.output.dense.bias.data,
                ln2_g=l.output.LayerNorm.weight.data, ln2_b=l.output.LayerNorm.bias.data))
        ew.wp = ops.cast_bf16(bert.pooler.dense.weight.data, out=old.wp if old is not None else None)
        ew.bp = bert.pooler.dense.bias.data
        return ew

    def _build_head_weights(self, old: Optional[engine.HeadWeights]):
        pr = self.cls.predictions
        return engine.HeadWeights(
            wt=ops.cast_bf16(pr.transform.dense.weight.data, out=old.wt if old else None),
            bt=pr.transform.dense.bias.data,
            ln_g=pr.transform.LayerNorm.weight.data, ln_b=pr.transform.LayerNorm.bias.data,
            w_text=ops.cast_bf16(pr.text_decoder.weight.data, out=old.w_text if old else None),
            w_ent=ops.cast_bf16(pr.entity_decoder.weight.data, out=old.w_ent if old else None),
            w_nsp=self.cls.seq_relationship.weight.data, b_nsp=self.cls.seq_relationship.bias.data)

    def _live_version(self):
        v = 0
        for p in self._enc_params(self.bert):
            v += p._version
        pr = self.cls.predictions
        for p in (pr.transform.dense.weight, pr.text_decoder.weight, pr.entity_decoder.weight):
            v += p._version
        return v

    def _lm_version(self):
        return sum(p._version for p in self.lm_backbone.parameters())

    def _device_state(self, need_heads: bool):
        """(Re)build the bf16 weight copies when parameters changed (optimizer step, load_state_dict)."""
        dev = self.bert.pooler.dense.weight.device
        if dev.type != "cuda":
            raise StkError("STonKGsForPreTraining runs on CUDA only: move the model with .to('cuda') "
                           "(stonkgs_b200 has no CPU path)")
        st = self._dev_state
        if st is None:
            st = self._dev_state = {"lm": None, "lm_v": None, "bert": None, "heads": None, "live_v": None,
                                    "heads_v": None}
        lm_v = self._lm_version()
        if st["lm"] is None or st["lm_v"] != lm_v:
            st["lm"] = self._build_encoder_weights(self.lm_backbone, True, st["lm"])
            st["lm_v"] = lm_v
            self._special_rows_version = None
        live_v = self._live_version()
        if st["bert"] is None or st["live_v"] != live_v:
            st["bert"] = self._build_encoder_weights(self.bert, False, st["bert"])
            st["live_v"] = live_v
        if need_heads and (st["heads"] is None or st["heads_v"] != live_v):
            st["heads"] = self._build_head_weights(st["heads"])
            st["heads_v"] = live_v
        if self._special_rows_version != lm_v:
            # rows 102/103/100 of the KG table are LM-backbone outputs of [[id]] (reference :138-141)
            if self.kg_table.device != dev:
                self.kg_table = self.kg_table.to(dev)
            ids = [i for i in (self.lm_sep_id, self.lm_mask_id, self.lm_unk_id) if i < self.kg_table.shape[0]]
            if ids:
                self.kg_table[torch.tensor(ids, device=dev)] = engine.lm_special_rows(st["lm"], ids)
            self._special_rows_version = lm_v
        return st

    # ------------------------------------------------------------------------------------------
    # forward
    # ------------------------------------------------------------------------------------------
    def _check_ids(self, input_ids):
        """The reference raises KeyError for ids outside the KG dict (stonkgs_model.py:182-189)."""
        if not input_ids.is_cuda:
            kg = input_ids[:, self.seq_shape.text_len:]
            if kg.numel() and (int(kg.min()) < 0 or int(kg.max()) >= self.kg_table.shape[0]):
                bad = kg[(kg < 0) | (kg >= self.kg_table.shape[0])][0]
                raise KeyError(int(bad))

    def encode(self, input_ids, attention_mask=None, token_type_ids=None, *, cache=None, want_inputs_embeds=False,
               need_heads=False, err_flag=None, head_mask=None, pooled_only=False, live_plan=None):
        """LM backbone -> KG lookup -> joint encoder -> pooler.  Returns (seq bf16 [B*seq_pad,768], pooled fp32, emb);
        with ``pooled_only`` (eval) the last layer runs on the [CLS] rows alone and ``seq`` is None; ``live_plan``:
        see engine.joint_fwd."""
        st = self._device_state(need_heads)
        dev = self.kg_table.device
        sh = self.seq_shape
        if input_ids.dim() != 2 or input_ids.shape[1] != sh.seq_len:
            raise StkError(f"input_ids must be [B, {sh.seq_len}] ({sh.text_len} text + {sh.kg_len} KG tokens), "
                           f"got {tuple(input_ids.shape)}")
        self._check_ids(input_ids)
        # ids already on the device are range-checked by the kernels (device flag, read later: _raise_on_bad_ids);
        # a caller streaming many batches passes ONE flag for all of them (embeddings.EmbeddingStreamer)
        err = err_flag if err_flag is not None else (torch.zeros(1, dtype=torch.int32, device=dev) if input_ids.is_cuda else None)
        input_ids = input_ids.to(dev, torch.int64, non_blocking=True).contiguous()
        if attention_mask is not None:
            attention_mask = attention_mask.to(dev, torch.int64, non_blocking=True).contiguous()
        if token_type_ids is not None:
            token_type_ids = token_type_ids.to(dev, torch.int64, non_blocking=True).contiguous()
        hm = self._head_mask_rows(head_mask, len(st["bert"].layers), dev)
        drop = self._drop_ctx()
        lm_hidden = engine.lm_backbone_fwd(st["lm"], input_ids[:, :sh.text_len], None, err_flag=err, drop=drop)
        seq, pooled, emb = engine.joint_fwd(st["bert"], input_ids, token_type_ids, attention_mask, lm_hidden,
                                            self.kg_table, cache=cache, want_inputs_embeds=want_inputs_embeds,
                                            err_flag=err, drop=drop, shape=sh, head_mask=hm,
                                            pooled_only=pooled_only and cache is None,
                                            live_plan=live_plan if cache is None else None)
        if cache is not None:
            cache.update(input_ids=input_ids, token_type_ids=token_type_ids, err=err)
        if err_flag is None:
            self._pending_err = err
        return seq, pooled, emb

    @staticmethod
    def _head_mask_rows(head_mask, num_layers: int, dev):
        """The reference hands ``head_mask`` to ``self.bert`` (stonkgs_model.py:158,209); HF's ``get_head_mask`` accepts
        [heads] (same for every layer) or [layers, heads] and multiplies the attention probabilities of (layer, head) by
        the entry.  Returns fp32 [layers, 12] on the device, or None."""
        if head_mask is None:
            return None
        hm = torch.as_tensor(head_mask, dtype=torch.float32).to(dev)
        if hm.dim() == 1:
            hm = hm.unsqueeze(0).expand(num_layers, -1)
        if hm.dim() != 2 or hm.shape != (num_layers, ops.HEADS):
            raise StkError(f"head_mask must be [12] or [{num_layers}, 12], got {tuple(hm.shape)}")
        return hm.contiguous()

    def _drop_ctx(self):
        """Dropout state of this forward pass: None in eval() (the reference's parity mode) or when disabled."""
        p_h = float(getattr(self.config, "hidden_dropout_prob", 0.0) or 0.0)
        p_a = float(getattr(self.config, "attention_probs_dropout_prob", 0.0) or 0.0)
        if not (self.training and getattr(self, "stk_dropout", False) and (p_h > 0.0 or p_a > 0.0)):
            self.last_dropout_seed = None
            return None
        seed = (self.stk_dropout_seed + 0x9E3779B9 * self._dropout_step) & 0xFFFFFFFF
        self._dropout_step += 1
        self.last_dropout_seed = seed
        return engine.DropCtx(seed, p_h, p_a)

    def grad_buffer(self):
        """Flat fp32 gradient buffer of the live parameters (``param.grad`` are views of it)."""
        from .training import GradBuffer
        gb = getattr(self, "_grad_buffer", None)
        dev = self.bert.pooler.dense.weight.device
        if gb is None or gb.flat.device != dev or gb.stale():
            gb = self._grad_buffer = GradBuffer(self)
        return gb

    def _stage_err_flag(self):
        """Training step: copy the id / label range flag of the forward just enqueued to pinned host memory on a side
        stream.  By the time ``FusedAdamW.step`` (or the next forward) looks at it, the copy has long completed behind the
        backward pass, so the check costs no device sync and never drains the launch queue."""
        err = getattr(self, "_pending_err", None)
        if err is None:
            return
        if getattr(self, "_err_stream", None) is None or self._err_host.device != torch.device("cpu"):
            self._err_stream = torch.cuda.Stream(device=err.device)
            self._err_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        fwd_done = torch.cuda.Event()
        fwd_done.record(torch.cuda.current_stream(err.device))
        with torch.cuda.stream(self._err_stream):
            self._err_stream.wait_event(fwd_done)
            self._err_host.copy_(err, non_blocking=True)
            err.record_stream(self._err_stream)
            self._err_staged = torch.cuda.Event()
            self._err_staged.record(self._err_stream)

    def _raise_on_bad_ids(self):
        """Read the device-side id / label range flag of the last forward and raise like the reference (KeyError from
        ``kg_backbone[i]``, IndexError from the embedding / cross-entropy).  Inference reads it right away (one sync);
        the training step stages it (``_stage_err_flag``) and checks in ``FusedAdamW.step`` before the update is applied,
        or at the next forward."""
        err = getattr(self, "_pending_err", None)
        staged = getattr(self, "_err_staged", None)
        self._pending_err = None
        self._err_staged = None
        if err is None:
            return
        if staged is not None:
            staged.synchronize()
            code = int(self._err_host[0])
        else:
            code = int(err.item())
        if code & 1:
            raise KeyError("input id outside the text vocabulary / KG table")
        if code & 2:
            raise IndexError("label outside the vocabulary of its head (text / entity / next-sentence)")
        if code & 4:
            raise StkError("more labelled positions in the batch than the head kernels were sized for: set "
                           "model.label_capacity (labelled positions per pair and half; default int(0.15 * 256) = 38 as "
                           "produced by the reference's pre-processing) or pass the label tensors on the CPU")

    @torch.no_grad()
    def embed(self, input_ids, attention_mask=None, token_type_ids=None, err_flag=None, pooling: str = "pooler",
              cls_rows_only: bool = False, skip_padding: bool = False, host_mask=None) -> torch.Tensor:
        """Extraction path, heads skipped.  ``pooling="pooler"`` (default) is the reference's output: the BERT pooler
        ``tanh(W h[:, 0] + b)`` (stonkgs_for_embeddings.py:180).  ``pooling="mean"`` is an extra: the mean of the last
        hidden state over the attended tokens (``attention_mask != 0``) of each pair.

        ``cls_rows_only`` (with the pooler, eval): the pooler reads ``hidden[:, 0]`` alone, so the last encoder layer is
        evaluated for the [CLS] rows only (keys / values still come from all 512 rows) — identical output, about 1/12 of
        the joint encoder's work less.  Off by default: every number quoted for this path runs the full last layer.

        ``skip_padding`` (with the pooler, eval, STonKGs shape): the joint encoder leaves out the padded rows of every
        pair — they are masked as keys and, as queries, only produce rows the pooler never reads — by packing each
        pair's attended rows first and running pairs that fit in 384 (256, 128) rows at that length
        (engine.plan_live_rows).  Same mathematics; the attention sums run over the keys in a different grouping, so the
        embeddings agree with the full pass to rounding (not bit for bit), except for pairs that keep all 512 rows.
        The plan is made on the host from ``host_mask`` (a host copy of ``attention_mask``; without it a device mask
        is copied back, which synchronises).  Off by default and never part of a quoted number."""
        if pooling not in ("pooler", "mean"):
            raise StkError(f"pooling must be 'pooler' or 'mean', got {pooling!r}")
        plan = None
        sh = self.seq_shape
        if (skip_padding and pooling == "pooler" and not self.training and attention_mask is not None
                and sh.seq_pad == sh.seq_len):
            hm = host_mask if host_mask is not None else attention_mask
            if isinstance(hm, torch.Tensor):
                hm = hm.cpu().numpy()
            plan = engine.plan_live_rows(hm)
        seq, pooled, _ = self.encode(input_ids, attention_mask, token_type_ids, err_flag=err_flag,
                                     pooled_only=cls_rows_only and pooling == "pooler" and not self.training,
                                     live_plan=plan)
        if pooling == "pooler":
            return pooled
        am = attention_mask
        if am is not None:
            am = am.to(seq.device, torch.int64, non_blocking=True).contiguous()
        return ops.masked_mean_pool(seq, am, input_ids.shape[0], self.seq_shape)

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, masked_lm_labels=None,
                ent_masked_lm_labels=None, next_sentence_labels=None, return_dict=None, head_mask=None):
        """Same contract as the reference forward (stonkgs_model.py:149-258)."""
        from . import training  # local import: keeps inference-only users free of the autograd glue
        return training.forward(self, input_ids, attention_mask, token_type_ids, masked_lm_labels,
                                ent_masked_lm_labels, next_sentence_labels, return_dict, head_mask)


# --------------------------------------------------------------------------------------------------
# TransE variant: same model on a 256 + 4 token sequence
# --------------------------------------------------------------------------------------------------
class TransESTonKGsELMPredictionHead(STonKGsELMPredictionHead):
    """Parameter layout of the reference head (transestonkgs_model.py:29-52): the text part is
    ``max_position_embeddings - 4`` positions long, the remaining four are entity positions."""

    def __init__(self, config):
        super().__init__(config)
        del self.half_length
        self.text_part_length = config.max_position_embeddings - 4


class TransESTonKGsForPreTraining(STonKGsForPreTraining):
    """Drop-in for ``TransESTonKGsForPreTraining`` (transestonkgs_model.py:70-250): 256 text tokens through the frozen
    LM backbone followed by 4 KG ids looked up in the TransE embedding table, ``max_position_embeddings`` = 260
    (:93).  Same kernels as STonKGs: the joint encoder runs on 384 rows per pair (three 128-key attention blocks) whose
    last 124 rows are zero and masked out as keys; outputs are sliced back to 260 positions."""

    seq_shape = ops.SeqShape(256, 260, 384)
    _head_class = TransESTonKGsELMPredictionHead

    @staticmethod
    def _adjust_config(config):
        config.update({"max_position_embeddings": 260})

    @classmethod
    @lru_cache(maxsize=32)
    def from_default_pretrained(cls, **kwargs):  # the reference publishes no TransE checkpoint
        raise StkError("there is no default pre-trained TransESTonKGs checkpoint")
