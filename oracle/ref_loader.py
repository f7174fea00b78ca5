"""Imports the UNMODIFIED reference hot-path modules from /root/reference under the
shim packages in oracle/ref_shims (detectron2 / fvcore / timm stand-ins).

TEST INFRASTRUCTURE.  Works only in the build container (the GPU box has no
/root/reference); used to (1) validate oracle/rba_oracle.py against the reference's
own forward and (2) generate the golden fixtures under tests/golden/ (see
oracle/make_golden.py).  Nothing in the product path, the `-m gpu` tests, smoke() or
bench.py imports this file.

Recipe follows SURVEY.md Appendix B: placeholder packages with __path__ into the
reference tree bypass the heavy mask2former/__init__.py (data mappers, evaluators);
a stub `MultiScaleDeformableAttention` module makes the reference's MSDeformAttn fall
back to its own pure-PyTorch statement `ms_deform_attn_core_pytorch`
(mask2former/modeling/pixel_decoder/ops/modules/ms_deform_attn.py:116-121).
"""
import importlib
import importlib.util
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("RBA_REFERENCE_ROOT", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mask2former"))


def _install():
    if "mask2former.maskformer_model" in sys.modules:
        return
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    msda = types.ModuleType("MultiScaleDeformableAttention")

    def _no_native(*a, **k):
        raise RuntimeError("reference CUDA op not built; reference falls back to its PyTorch statement")

    msda.ms_deform_attn_forward = _no_native
    msda.ms_deform_attn_backward = _no_native
    sys.modules["MultiScaleDeformableAttention"] = msda
    for name, rel in [
        ("mask2former", "mask2former"),
        ("mask2former.modeling", "mask2former/modeling"),
        ("mask2former.modeling.backbone", "mask2former/modeling/backbone"),
        ("mask2former.modeling.meta_arch", "mask2former/modeling/meta_arch"),
        ("mask2former.modeling.transformer_decoder", "mask2former/modeling/transformer_decoder"),
        ("mask2former.modeling.pixel_decoder", "mask2former/modeling/pixel_decoder"),
        ("mask2former.utils", "mask2former/utils"),
    ]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF_ROOT, rel)]
        sys.modules[name] = m
    for mod in [
        "mask2former.modeling.backbone.swin",
        "mask2former.modeling.pixel_decoder.msdeformattn",
        "mask2former.modeling.transformer_decoder.mask2former_transformer_decoder",
        "mask2former.modeling.meta_arch.mask_former_head",
        "mask2former.maskformer_model",
    ]:
        importlib.import_module(mod)


def _merge(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = v


def load_cfg(name_or_path="swin_b_1dl", overrides=None):
    """Dumped ckpt YAMLs are plain YAML; returns a shim CfgNode."""
    _install()
    from detectron2.config import CfgNode

    path = name_or_path
    if not os.path.isfile(path):
        path = os.path.join(REF_ROOT, "ckpts", name_or_path, "config.yaml")
    # defaults exactly as train_net.setup builds them (train_net.py:356-362): the reference's own
    # add_maskformer2_config over a skeleton of the detectron2 nodes it touches, then the YAML on top.
    spec = importlib.util.spec_from_file_location("_ref_m2f_config", os.path.join(REF_ROOT, "mask2former", "config.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = CfgNode({"INPUT": {"CROP": {}}, "SOLVER": {}, "MODEL": {"SEM_SEG_HEAD": {}}, "DATASETS": {}, "TEST": {}})
    mod.add_maskformer2_config(cfg)
    with open(path) as f:
        _merge(cfg, CfgNode(yaml.safe_load(f)))
    for dotted, v in (overrides or {}).items():
        node = cfg
        keys = dotted.split(".")
        for k in keys[:-1]:
            node = node[k]
        node[keys[-1]] = v
    cfg.MODEL.DEVICE = "cpu"
    return cfg


def build_reference_model(cfg, seed=0):
    """MaskFormer(cfg) exactly as detectron2's build_model would construct it."""
    _install()
    from mask2former.maskformer_model import MaskFormer

    torch.manual_seed(seed)
    model = MaskFormer(cfg)
    model.eval()
    return model


def msda_core_pytorch():
    _install()
    from mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

    return ms_deform_attn_core_pytorch
