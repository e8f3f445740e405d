"""stonkgs_b200 — B200-native implementation of the STonKGs joint text-KG transformer hot path.

Public surface (same names as the reference package ``stonkgs``):

* ``STonKGsForPreTraining``, ``STonKGsELMPredictionHead``, ``BertForPreTrainingOutputWithPooling``
  (reference ``stonkgs/models/stonkgs_model.py``)
* ``get_stonkgs_embeddings`` (reference ``stonkgs/models/stonkgs_for_embeddings.py:158-186``)
* ``TransESTonKGsForPreTraining``, ``TransESTonKGsELMPredictionHead`` (reference
  ``stonkgs/models/transestonkgs_model.py``): the 256 + 4 token variant of the same model

All compute runs in ``libstk.so`` (hand-written sm_100a CUDA, C ABI in ``include/stk.h``); importing
the model classes requires ``transformers`` only as the container of config / checkpoint layout.
"""
from ._lib import StkError, load as load_library  # noqa: F401

__all__ = ["StkError", "load_library", "STonKGsForPreTraining", "STonKGsELMPredictionHead",
           "BertForPreTrainingOutputWithPooling", "get_stonkgs_embeddings", "TransESTonKGsForPreTraining",
           "TransESTonKGsELMPredictionHead"]


def __getattr__(name):  # lazy: keep `import stonkgs_b200` cheap (no transformers import)
    if name in ("STonKGsForPreTraining", "STonKGsELMPredictionHead", "BertForPreTrainingOutputWithPooling",
                "TransESTonKGsForPreTraining", "TransESTonKGsELMPredictionHead"):
        from . import model
        return getattr(model, name)
    if name == "get_stonkgs_embeddings":
        from .embeddings import get_stonkgs_embeddings
        return get_stonkgs_embeddings
    raise AttributeError(name)
