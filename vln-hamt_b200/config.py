"""Model configuration for the HAMT hot path.

Mirrors the keys of the reference's HF-style model config JSON
(pretrain_src/config/r2r_model_config.json:1-33, rxr_xlm_model_config.json) plus the runtime-injected
ones: ``pretrain_tasks`` (pretrain_src/main_r2r.py:123-126) and the finetune flags
(finetune_src/models/vlnbert_init.py:37-63).  Any object with these attributes (e.g. a
``transformers.PretrainedConfig``) is accepted by the modules; this class exists so the package does
not depend on ``transformers``.
"""
from __future__ import annotations

import copy
import json

R2R_MODEL_CONFIG = {
    "pred_head_dropout_prob": 0.1,
    "attention_probs_dropout_prob": 0.1,
    "hidden_act": "gelu",
    "hidden_dropout_prob": 0.1,
    "hidden_size": 768,
    "image_feat_size": 768,
    "angle_feat_size": 4,
    "image_prob_size": 1000,
    "img_feature_type": "imagenet",
    "initializer_range": 0.02,
    "intermediate_size": 3072,
    "num_l_layers": 9,
    "num_r_layers": 0,
    "num_h_layers": 0,
    "num_x_layers": 4,
    "num_h_pano_layers": 2,
    "layer_norm_eps": 1e-12,
    "max_position_embeddings": 512,
    "max_action_steps": 100,
    "num_attention_heads": 12,
    "num_hidden_layers": 12,
    "output_attentions": False,
    "output_hidden_states": False,
    "type_vocab_size": 2,
    "update_lang_bert": True,
    "vocab_size": 30522,
    "lang_bert_name": "bert-base-uncased",
}

# rxr_xlm_model_config.json differences
RXR_OVERRIDES = {"image_feat_size": 512, "vocab_size": 250002, "max_position_embeddings": 514,
                 "lang_bert_name": "xlm-roberta-base"}

ALL_TASKS = ("mlm", "sap", "sar", "sprel", "mrc", "itm")


class HamtConfig:
    def __init__(self, **kw):
        d = dict(R2R_MODEL_CONFIG)
        d.update(kw)
        for k, v in d.items():
            setattr(self, k, v)
        if not hasattr(self, "pretrain_tasks"):
            self.pretrain_tasks = set(ALL_TASKS)
        # finetune flags default to the shipped R2R recipe's semantics (vlnbert_init.py:49-61)
        for k, v in dict(hist_enc_pano=True, fix_lang_embedding=False, fix_hist_embedding=False,
                         fix_obs_embedding=False, no_lang_ca=False, act_pred_token="ob_txt").items():
            if not hasattr(self, k):
                setattr(self, k, v)

    @classmethod
    def from_json_file(cls, path, **kw):
        with open(path) as f:
            d = json.load(f)
        d.update(kw)
        return cls(**d)

    @classmethod
    def rxr(cls, **kw):
        d = dict(RXR_OVERRIDES)
        d.update(kw)
        return cls(**d)

    def to_dict(self):
        d = dict(self.__dict__)
        d["pretrain_tasks"] = sorted(d["pretrain_tasks"])
        return d

    def __copy__(self):
        c = HamtConfig.__new__(HamtConfig)
        c.__dict__.update(self.__dict__)
        return c

    def copy(self):
        return copy.copy(self)
