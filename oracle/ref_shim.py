"""TEST INFRASTRUCTURE — side-effect-free import of the real reference module.

Executes the reference's own ``stonkgs.models.stonkgs_model.STonKGsForPreTraining`` on CPU, from
``/root/reference/src`` (dev container) or from ``baseline/_ref`` (the UNMODIFIED reference package installed with
``pip install --no-index --no-deps --ignore-requires-python --target baseline/_ref <copy of /root/reference>``; git-ignored,
it travels to the GPU box, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg time it), following the
recipe that SURVEY.md Appendix A verified:

* stub ``mlflow`` / ``pytorch_lightning`` (imported by ``kg_baseline_model.py:16,19`` which
  ``stonkgs_model.py:23`` pulls in only for ``prepare_df``);
* a namespace stub for the ``stonkgs`` package so ``stonkgs/__init__.py`` (-> api -> pybel/indra)
  is not executed, and a stub for ``stonkgs.constants`` (the real one downloads a vocabulary and
  ``os.makedirs`` inside the reference tree at import, ``constants.py:91-110,129``);
* offline replacements of the three hub-bound constructors
  (``stonkgs_model.py:96,107,116-118``);
* ``prepare_df`` patched to return the synthetic node2vec rows as float64, the dtype the TSV
  parser yields (``kg_baseline_model.py:270-280``).

Only ``tests/`` (CPU, dev container), ``oracle/make_golden.py`` and the two CPU-baseline legs of ``bench.py`` import
this file; the product never does.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CANDIDATES = ("/root/reference/src", os.path.join(_REPO, "baseline", "_ref"))


def _find_reference():
    for c in _CANDIDATES:
        if os.path.isfile(os.path.join(c, "stonkgs", "models", "stonkgs_model.py")):
            return c
    return None


REFERENCE_SRC = _find_reference() or _CANDIDATES[0]


def reference_available() -> bool:
    return _find_reference() is not None


_state = {"sm": None, "num_layers": 12, "lm_sd": None}


def _import_reference():
    if _state["sm"] is not None:
        return _state["sm"]
    if not reference_available():
        raise RuntimeError("reference tree not present (dev container only)")
    sys.dont_write_bytecode = True
    sys.modules.setdefault("mlflow", types.ModuleType("mlflow"))
    if "pytorch_lightning" not in sys.modules:
        pl = types.ModuleType("pytorch_lightning")
        pl.LightningModule = torch.nn.Module
        pl.Trainer = object
        sys.modules["pytorch_lightning"] = pl
    pkg = types.ModuleType("stonkgs")
    pkg.__path__ = [os.path.join(REFERENCE_SRC, "stonkgs")]
    sys.modules["stonkgs"] = pkg

    class _C(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith("__"):
                raise AttributeError(name)
            return "/nonexistent/" + name

    c = _C("stonkgs.constants")
    c.NLP_MODEL_TYPE = "dmis-lab/biobert-v1.1"
    c.EMBEDDINGS_PATH = "/nonexistent/emb.tsv"
    sys.modules["stonkgs.constants"] = c
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)

    import stonkgs.models.stonkgs_model as sm  # noqa: E402  (the reference's own module)

    _state["sm"] = sm
    return sm


class _offline_hub:
    """While active, the three hub-bound constructors the reference calls in ``__init__`` (stonkgs_model.py:96,107,
    116-118) resolve to offline equivalents (BioBERT-v1.1 shape, weights from ``_state``); restored on exit so that
    other code in the same process (e.g. a save_pretrained / from_pretrained round trip) sees the real ones."""

    def __enter__(self):
        import transformers
        from transformers import BertConfig, BertModel
        self._saved = (BertConfig.__dict__.get("from_pretrained"), BertModel.__dict__.get("from_pretrained"),
                       transformers.BertTokenizer.__dict__.get("from_pretrained"))

        def _cfg():
            cfg = BertConfig(vocab_size=28996, num_hidden_layers=_state["num_layers"])
            cfg._attn_implementation = "eager"  # the 2021 code path; canonical oracle
            return cfg

        def _bert_from_pretrained(cls, *a, **k):
            m = BertModel(_cfg()).eval()
            if _state["lm_sd"] is not None:
                m.load_state_dict(_state["lm_sd"], strict=True)
            return m

        class _Tok:
            sep_token_id, mask_token_id, unk_token_id = 102, 103, 100

        BertConfig.from_pretrained = classmethod(lambda cls, *a, **k: _cfg())
        BertModel.from_pretrained = classmethod(_bert_from_pretrained)
        transformers.BertTokenizer.from_pretrained = classmethod(lambda cls, *a, **k: _Tok())
        return self

    def __exit__(self, *exc):
        import transformers
        from transformers import BertConfig, BertModel
        for cls, saved in zip((BertConfig, BertModel, transformers.BertTokenizer), self._saved):
            if saved is None:
                try:
                    delattr(cls, "from_pretrained")     # fall back to the inherited classmethod
                except AttributeError:
                    pass
            else:
                setattr(cls, "from_pretrained", saved)
        return False


def load_reference(state_dict, kg_table: np.ndarray, num_layers: int = 12):
    """Instantiate the reference model with the given synthetic checkpoint and node2vec rows."""
    sm = _import_reference()
    _state["num_layers"] = num_layers
    # the LM backbone must hold its final weights *during* __init__ because the reference computes
    # the [SEP]/[MASK]/[UNK] rows of kg_backbone there (stonkgs_model.py:138-141)
    _state["lm_sd"] = {
        k[len("lm_backbone."):]: v for k, v in state_dict.items() if k.startswith("lm_backbone.")
    }
    tab64 = kg_table.astype(np.float64)
    sm.prepare_df = lambda path: {f"n{i}": tab64[i] for i in range(tab64.shape[0])}
    was_cuda = torch.cuda.is_available
    torch.cuda.is_available = lambda: False  # keep lm_backbone on CPU (stonkgs_model.py:109-110)
    try:
        with _offline_hub():
            model = sm.STonKGsForPreTraining(config=None, kg_embedding_dict_path="unused")
    finally:
        torch.cuda.is_available = was_cuda
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()


def load_reference_transe(state_dict, kg_table: np.ndarray, num_layers: int = 12):
    """The reference's ``TransESTonKGsForPreTraining`` (transestonkgs_model.py:70-250) with a synthetic checkpoint
    (``bert.embeddings.position_embeddings`` is [260, 768]) and synthetic TransE rows."""
    _import_reference()
    import stonkgs.models.transestonkgs_model as tm  # noqa: E402  (the reference's own module)
    _state["num_layers"] = num_layers
    _state["lm_sd"] = {
        k[len("lm_backbone."):]: v for k, v in state_dict.items() if k.startswith("lm_backbone.")
    }
    tab64 = kg_table.astype(np.float64)
    tm.prepare_df = lambda path: {f"n{i}": tab64[i] for i in range(tab64.shape[0])}
    was_cuda = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with _offline_hub():
            model = tm.TransESTonKGsForPreTraining(config=None, kg_embedding_dict_path="unused")
    finally:
        torch.cuda.is_available = was_cuda
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()


def load_reference_classifier(state_dict, kg_table: np.ndarray, num_layers: int, num_labels: int):
    """The reference's ``STonKGsForSequenceClassification`` (stonkgs_finetuning.py:237-346) with the given
    synthetic checkpoint (pre-training keys + ``classifier.*``)."""
    sm = _import_reference()
    import stonkgs.models.stonkgs_finetuning as sf  # noqa: E402  (the reference's own module)
    from transformers import BertConfig
    _state["num_layers"] = num_layers
    _state["lm_sd"] = {k[len("lm_backbone."):]: v for k, v in state_dict.items() if k.startswith("lm_backbone.")}
    tab64 = kg_table.astype(np.float64)
    sm.prepare_df = lambda path: {f"n{i}": tab64[i] for i in range(tab64.shape[0])}
    cfg = BertConfig(vocab_size=28996, num_hidden_layers=num_layers, num_labels=num_labels)
    cfg._attn_implementation = "eager"
    cfg.kg_vocab_size = kg_table.shape[0]
    was_cuda = torch.cuda.is_available
    torch.cuda.is_available = lambda: False
    try:
        with _offline_hub():
            model = sf.STonKGsForSequenceClassification(cfg, kg_embedding_dict_path="unused")
    finally:
        torch.cuda.is_available = was_cuda
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    return model.eval()
