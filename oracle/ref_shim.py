"""Import the UNMODIFIED reference modules (cshizhe/VLN-HAMT) under torch 2.11 / transformers 5.x.

TEST / BASELINE INFRASTRUCTURE ONLY.  Needs the reference's Python sources: the checkout in the build container
(/root/reference) or the verbatim copy `__graft_entry__.build()` installs under baseline/_ref (git-ignored; it
travels to the GPU box with the snapshot, where it is what `bench.py --impl reference` times).  The GPU tests never
import it: they use the committed golden vectors produced through this shim by oracle/make_golden.py.

The shim follows SURVEY.md section 8c: the reference pins transformers==4.12.3
(requirements.txt:11) and uses three things from it that transformers 5 removed/changed:
  * transformers.modeling_utils.get_parameter_device            (vilmodel.py:16)
  * BertPreTrainedModel.init_weights / _tie_or_clone_weights    (vilmodel.py:589, pretrain_cmt.py:93-99)
  * from_pretrained(None, config=..., state_dict=...)           (main_r2r.py:146-148)
None of these contributes forward arithmetic.  No reference file is edited or copied.
"""
from __future__ import annotations

import importlib
import os
import sys

import torch
import torch.nn as nn

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _resolve_root() -> str:
    """The reference checkout (build container) or the verbatim copy of its Python sources that `__graft_entry__.build()` installs
    under baseline/_ref (git-ignored; travels to the GPU box like the built .so)."""
    for cand in (os.environ.get("HAMT_REFERENCE_ROOT"), "/root/reference", os.path.join(_REPO_ROOT, "baseline", "_ref")):
        if cand and os.path.isfile(os.path.join(cand, "pretrain_src", "model", "vilmodel.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _resolve_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "pretrain_src", "model", "vilmodel.py"))


class _CompatBertPreTrainedModel(nn.Module):
    """Stands in for transformers==4.12.3 BertPreTrainedModel (init / tying / (de)serialisation only)."""
    base_model_prefix = "bert"

    def __init__(self, config, *a, **k):
        super().__init__()
        self.config = config

    def _init_weights(self, mod):
        if isinstance(mod, (nn.Linear, nn.Embedding)):
            mod.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(mod, nn.LayerNorm):
            mod.bias.data.zero_()
            mod.weight.data.fill_(1.0)
        if isinstance(mod, nn.Linear) and mod.bias is not None:
            mod.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)

    def _tie_or_clone_weights(self, a, b):
        a.weight = b.weight

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path=None, config=None, state_dict=None, **kw):
        model = cls(config)
        if state_dict:
            model.load_state_dict(state_dict, strict=False)
        return model


_LOADED = {}


def _patch_transformers():
    import transformers
    import transformers.modeling_utils as mu
    mu.get_parameter_device = lambda m: next(m.parameters()).device
    transformers.BertPreTrainedModel = _CompatBertPreTrainedModel


def _import_from(subdir: str, modname: str):
    """Import `modname` with <reference>/<subdir> at the front of sys.path, isolated from the other
    tree (pretrain_src and finetune_src both have top-level packages named `utils`)."""
    key = (subdir, modname)
    if key in _LOADED:
        return _LOADED[key]
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    _patch_transformers()
    sys.dont_write_bytecode = True          # /root/reference is read-only
    path = os.path.join(REFERENCE_ROOT, subdir)
    # purge clashing top-level packages from a previous import of the other tree
    for name in list(sys.modules):
        if name.split(".")[0] in ("model", "models", "utils", "data", "optim"):
            f = getattr(sys.modules[name], "__file__", "") or ""
            if REFERENCE_ROOT in f and path not in f:
                del sys.modules[name]
    sys.path.insert(0, path)
    try:
        mod = importlib.import_module(modname)
    finally:
        sys.path.remove(path)
    _LOADED[key] = mod
    return mod


def pretrain_config(tasks=("mlm", "sap", "sar", "sprel", "mrc", "itm"), config_name="r2r_model_config.json", **overrides):
    from transformers import PretrainedConfig
    cfg = PretrainedConfig.from_json_file(os.path.join(REFERENCE_ROOT, "pretrain_src", "config", config_name))
    cfg.pretrain_tasks = set(tasks)                 # main_r2r.py:123-126
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def load_pretrain_model(cfg):
    """-> reference MultiStepNavCMTPreTraining (pretrain_src/model/pretrain_cmt.py:73)."""
    mod = _import_from("pretrain_src", "model.pretrain_cmt")
    return mod.MultiStepNavCMTPreTraining(cfg)


def load_navcmt(cfg):
    """-> reference NavCMT (finetune_src/models/vilmodel_cmt.py:610)."""
    mod = _import_from("finetune_src", "models.vilmodel_cmt")
    return mod.NavCMT(cfg)


# --------------------------------------------------------------------------------------------------------------------------
# end-to-end stage: the reference's vision backbone is a vendored timm file that imports five helper names from
# timm (vision_transformer.py:30-33); timm is not installed here.  The stand-ins below let the UNMODIFIED file import; none of
# them contributes forward arithmetic (DropPath is only instantiated for drop_path > 0, the reference passes 0:
# image_vilmodel.py:26-29; the init helpers only draw initial weights, which the tests overwrite with a seeded state_dict).
# --------------------------------------------------------------------------------------------------------------------------
def _install_timm_stubs():
    import types
    if "timm" in sys.modules and not getattr(sys.modules["timm"], "_hamt_stub", False):
        return                                  # a real timm is importable: use it
    def mod(name):
        m = types.ModuleType(name)
        m._hamt_stub = True
        sys.modules[name] = m
        return m
    timm, data, models = mod("timm"), mod("timm.data"), mod("timm.models")
    helpers, layers, registry = mod("timm.models.helpers"), mod("timm.models.layers"), mod("timm.models.registry")
    timm.data, timm.models = data, models
    models.helpers, models.layers, models.registry = helpers, layers, registry
    data.IMAGENET_DEFAULT_MEAN, data.IMAGENET_DEFAULT_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
    helpers.build_model_with_cfg = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("timm stub: build the class directly"))
    helpers.overlay_external_default_cfg = lambda cfg, kw: cfg

    class DropPath(nn.Module):                  # never instantiated with drop_path = 0 (vision_transformer.py:190)
        def __init__(self, p=0.0):
            super().__init__()
            if p:
                raise RuntimeError("timm stub: stochastic depth is not part of the HAMT path (drop_path_rate = 0)")

        def forward(self, x):
            return x

    layers.DropPath = DropPath
    layers.to_2tuple = lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x)
    layers.trunc_normal_ = lambda t, std=1.0, **k: nn.init.trunc_normal_(t, std=std, a=-2 * std, b=2 * std)
    layers.lecun_normal_ = lambda t: nn.init.normal_(t, std=(1.0 / t.shape[1]) ** 0.5 if t.dim() > 1 else 1.0)
    registry.register_model = lambda fn: fn


def load_reference_vit(depth: int = 12, drop_rate: float = 0.0, attn_drop_rate: float = 0.0, **kw):
    """-> the reference's VisionTransformer (pretrain_src/model/vision_transformer.py:226) at ViT-B/16 geometry
    (vit_base_patch16_224 :507-513: patch 16, embed 768, heads 12; `depth` may be reduced for small fixtures), num_classes = 0."""
    _install_timm_stubs()
    if not reference_available():
        raise RuntimeError(f"reference checkout not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    import importlib.util
    path = os.path.join(REFERENCE_ROOT, "pretrain_src", "model", "vision_transformer.py")
    key = ("file", path)
    if key not in _LOADED:
        spec = importlib.util.spec_from_file_location("hamt_ref_vision_transformer", path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _LOADED[key] = m
    vt = _LOADED[key]
    return vt.VisionTransformer(patch_size=16, embed_dim=768, depth=depth, num_heads=12, num_classes=0, drop_rate=drop_rate,
                                attn_drop_rate=attn_drop_rate, drop_path_rate=0.0, **kw)
