"""hamt_b200 -- Blackwell-native (sm_100a) implementation of the HAMT data-parallel hot path.

The directory is named ``vln-hamt_b200`` (not importable as such); import it as ``hamt_b200``
through the loader module ``hamt_b200.py`` at the repo root.

Reference-facing API (same names / signatures as cshizhe/VLN-HAMT):
  hamt_b200.pretrain_cmt.MultiStepNavCMTPreTraining   <- pretrain_src/model/pretrain_cmt.py:73
  hamt_b200.vilmodel.NavPreTrainedModel               <- pretrain_src/model/vilmodel.py:578
  hamt_b200.vilmodel_cmt.NavCMT                       <- finetune_src/models/vilmodel_cmt.py:610
  hamt_b200.model_HAMT.VLNBertCMT / Critic            <- finetune_src/models/model_HAMT.py:11,258
"""
from .config import HamtConfig, ALL_TASKS  # noqa: F401

__all__ = ["HamtConfig", "ALL_TASKS"]
__version__ = "0.1.0"
