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

_HERE = os.path.dirname(os.path.abspath(__file__))
_BASELINE_REF = os.path.join(os.path.dirname(_HERE), "baseline", "_ref")      # tools/make_baseline_ref.py (travels to the GPU box)
REF_ROOT = os.environ.get("RBA_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/mask2former") else _BASELINE_REF)
_SHIMS = os.path.join(_HERE, "ref_shims")
NATIVE_MSDA = False      # set by use_native_msda(): the reference's own CUDA extension instead of the raising stub


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "mask2former"))


def native_msda_available():
    import glob
    return bool(glob.glob(os.path.join(_BASELINE_REF, "MultiScaleDeformableAttention*.so")))


def use_native_msda():
    """Bind the reference's own MultiScaleDeformableAttention extension (built unmodified-kernels for sm_100a by
    tools/make_baseline_ref.py) instead of the stub.  Must be called before the first build/load."""
    global NATIVE_MSDA
    assert "mask2former.maskformer_model" not in sys.modules, "use_native_msda() must precede the first import"
    assert native_msda_available(), "baseline/_ref has no MultiScaleDeformableAttention*.so (run tools/make_baseline_ref.py)"
    NATIVE_MSDA = True


def _install():
    if "mask2former.maskformer_model" in sys.modules:
        return
    if _SHIMS not in sys.path:
        sys.path.insert(0, _SHIMS)
    if NATIVE_MSDA:
        if _BASELINE_REF not in sys.path:
            sys.path.insert(0, _BASELINE_REF)
        import MultiScaleDeformableAttention  # noqa: F401  (the real extension; registers itself in sys.modules)
    else:
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
    # register the build_resnet_backbone stand-in in whichever detectron2 stand-in is active (oracle/ref_shims, or the
    # product's rba_b200.compat stand-ins when a test plugged those in first)
    from detectron2.modeling import BACKBONE_REGISTRY
    try:
        BACKBONE_REGISTRY.get("build_resnet_backbone")
    except KeyError:
        spec = importlib.util.spec_from_file_location(
            "_ref_shim_resnet", os.path.join(_SHIMS, "detectron2", "modeling", "backbone", "resnet.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
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


def r50_overrides(dec_layers=1, levels=1, depth=50):
    """The R50 variants have no dumped ckpt YAML: derive them from the swin_b_1dl dump (every detectron2 default key is
    spelled out there, MODEL.RESNETS included) by switching the backbone, as configs/cityscapes/semantic-segmentation/
    maskformer2_R50_bs16_90k.yaml + Base-Cityscapes-SemanticSegmentation.yaml:4-15 do."""
    o = {"MODEL.BACKBONE.NAME": "build_resnet_backbone", "MODEL.RESNETS.DEPTH": depth, "MODEL.RESNETS.STRIDE_IN_1X1": False,
         "MODEL.RESNETS.OUT_FEATURES": ["res2", "res3", "res4", "res5"], "MODEL.MASK_FORMER.DEC_LAYERS": dec_layers + 1}
    if levels == 3:
        o["MODEL.SEM_SEG_HEAD.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES"] = ["res3", "res4", "res5"]
    return o


def build_reference_model(cfg, seed=0):
    """MaskFormer(cfg) exactly as detectron2's build_model would construct it."""
    _install()
    from mask2former.maskformer_model import MaskFormer

    torch.manual_seed(seed)
    model = MaskFormer(cfg)
    model.eval()
    return model


def load_state_dict_into(model, sd):
    """Loads a reference-layout state_dict keeping the module versions (no legacy-key upgrade fires)."""
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "state_dict keys differ from the reference"
    sd_meta = type(ref_sd)(sd)
    sd_meta._metadata = ref_sd._metadata
    model.load_state_dict(sd_meta)
    return model


def spy_attention_decisions(model):
    """Wraps the reference's forward_prediction_heads (mask2former_transformer_decoder.py:472-489) and returns the list
    its boolean attention masks are appended to, one (B,Q,S_l) bool tensor per head (head 0 of the nheads copies,
    BEFORE the all-blocked-row reset of :433)."""
    pred = model.sem_seg_head.predictor
    orig = pred.forward_prediction_heads
    am_list = []

    def spy(output, mask_features, attn_mask_target_size):
        r = orig(output, mask_features, attn_mask_target_size)
        am = r[2]
        B = mask_features.shape[0]
        am_list.append(am.view(B, -1, am.shape[1], am.shape[2])[:, 0].clone())
        return r

    pred.forward_prediction_heads = spy
    return am_list


def msda_core_pytorch():
    _install()
    from mask2former.modeling.pixel_decoder.ops.functions.ms_deform_attn_func import ms_deform_attn_core_pytorch

    return ms_deform_attn_core_pytorch
