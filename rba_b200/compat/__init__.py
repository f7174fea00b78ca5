"""Drop-in plumbing: lets the reference's own scripts (evaluate_ood.py, train_net.setup/Trainer.build_model) run
UNCHANGED with rba_b200 serving the `MaskFormer` meta-architecture.

    python -m rba_b200.compat.run evaluate_ood.py --models_folder ckpts/ --datasets_folder /data ...

`plug_in()` does three things, none of which touches the reference's files:
  1. `standins.install()`: detectron2 / fvcore / timm / easydict / albumentations / ... stand-ins for the packages
     that are not installed (a real install always wins);
  2. the top-level module `MultiScaleDeformableAttention` (the reference's only native FFI, ops/src/vision.cpp:18-21)
     resolves to rba_b200.compat.MultiScaleDeformableAttention -> rba_msda_forward;
  3. `build_model(cfg)` returns `rba_b200.MaskFormer` when cfg.MODEL.META_ARCHITECTURE == "MaskFormer", whatever the
     reference registered under that name.
"""
import importlib
import sys

from . import standins


def plug_in(meta_arch_names=("MaskFormer",)):
    served = standins.install()
    if "MultiScaleDeformableAttention" not in sys.modules:
        sys.modules["MultiScaleDeformableAttention"] = importlib.import_module("rba_b200.compat.MultiScaleDeformableAttention")
    from ..modeling import MaskFormer
    for n in meta_arch_names:
        standins.PLUGIN_META_ARCH[n] = MaskFormer
    if "detectron2" not in served:
        _patch_real_detectron2()
    return served


def _patch_real_detectron2():
    """With a real detectron2: wrap build_model so the plug-in table is consulted first (the reference's own class
    stays registered in META_ARCH_REGISTRY)."""
    import detectron2.modeling as dm
    from detectron2.modeling.meta_arch import build as dbuild
    import torch

    if getattr(dbuild.build_model, "_rba_b200", False):
        return
    orig = dbuild.build_model

    def build_model(cfg):
        cls = standins.PLUGIN_META_ARCH.get(cfg.MODEL.META_ARCHITECTURE)
        if cls is None:
            return orig(cfg)
        model = cls(cfg)
        model.to(torch.device(cfg.MODEL.DEVICE))
        return model

    build_model._rba_b200 = True
    dbuild.build_model = build_model
    dm.build_model = build_model
    import detectron2.modeling.meta_arch as dma
    dma.build_model = build_model
