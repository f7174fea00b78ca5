"""Stand-ins for the third-party packages the reference's `evaluate_ood.py` imports transitively (SURVEY.md §8b,
"compat surface"): detectron2, fvcore, timm, easydict, albumentations, matplotlib, ood_metrics, webp, fairscale,
panopticapi, pycocotools, shapely.  None of them is installed on the B200 image and none is on the arithmetic path of
the hot loop — the reference uses them for configuration, registries, checkpoint I/O and data transforms.

Two kinds of modules are provided, and only for packages that are NOT really importable (a real install always wins):
  * functional stand-ins for what the eval path actually executes (lenient yacs-style CfgNode + get_cfg /
    add_deeplab_config, `configurable`, registries, `build_model` -> rba_b200.MaskFormer, DetectionCheckpointer,
    MetadataCatalog / DatasetCatalog, default_setup, comm, ImageList, sem_seg_postprocess, EasyDict,
    albumentations.Compose / Resize / ToTensorV2, ood_metrics.fpr_at_95_tpr, ...), each citing the call site it serves;
  * inert placeholders for names that only have to import (training mappers, COCO/LVIS evaluators, TTA, ...): any
    attribute of such a module is a `Stub` that can be subclassed, called as a decorator or instantiated, and raises
    `RbaError` only if someone tries to compute with it.
"""
import copy
import importlib.abc
import importlib.machinery
import importlib.util
import logging
import os
import pickle
import sys
import types

import numpy as np
import torch
import yaml
from torch import nn
from torch.nn import functional as F

from .._lib import RbaError

# packages fully owned by the stand-in layer when the real one is absent
STANDIN_ROOTS = ("detectron2", "fvcore", "timm", "easydict", "albumentations", "matplotlib", "ood_metrics", "webp",
                 "fairscale", "panopticapi", "pycocotools", "shapely", "lvis", "cityscapesscripts")


# ----------------------------------------------------------------------------------------------------------------
# inert placeholders
# ----------------------------------------------------------------------------------------------------------------
class _StubMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return make_stub(f"{cls.__name__}.{name}")

    # a placeholder used as DATA at import time (e.g. builtin_meta.COCO_CATEGORIES) reads as an empty collection
    def __iter__(cls):
        return iter(())

    def __len__(cls):
        return 0

    def __contains__(cls, item):
        return False

    def __getitem__(cls, key):
        return make_stub(f"{cls.__name__}[{key!r}]")


class Stub(metaclass=_StubMeta):
    """Importable placeholder: subclassable, usable as `@decorator` / `@decorator(...)`, attribute access yields
    further stubs.  It stands for a third-party symbol the inference path never executes."""

    _stub_name = "stub"

    def __init__(self, *args, **kwargs):
        pass

    def __new__(cls, *args, **kwargs):
        # `@stub` applied to a function / class: pass it through unchanged
        if cls.__dict__.get("_is_leaf_stub", False) and len(args) == 1 and not kwargs and callable(args[0]) \
                and not isinstance(args[0], Stub):
            return args[0]
        return super().__new__(cls)

    def __call__(self, *args, **kwargs):
        if len(args) == 1 and not kwargs and callable(args[0]) and not isinstance(args[0], Stub):
            return args[0]                      # `@stub(...)` used as a decorator factory
        return self

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return make_stub(f"{type(self).__name__}.{name}")

    def __iter__(self):
        return iter(())

    def __bool__(self):
        return False

    def compute(self, *a, **k):
        raise RbaError(f"{type(self)._stub_name} is an import-only placeholder of rba_b200.compat (not on the inference path)")


def make_stub(name):
    return _StubMeta(name.split(".")[-1] or "Stub", (Stub,), {"_stub_name": name, "_is_leaf_stub": True})


class _StubModule(types.ModuleType):
    """Module whose every missing attribute is a placeholder (and which is also a package, so that
    `import a.b.c` works for any depth)."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        s = make_stub(f"{self.__name__}.{name}")
        setattr(self, name, s)
        return s


# ----------------------------------------------------------------------------------------------------------------
# detectron2.config (train_net.py:356-362, maskformer_model.py:29,107, mask2former/config.py)
# ----------------------------------------------------------------------------------------------------------------
class CfgNode(dict):
    """Lenient yacs-style node.  Differences from yacs, on purpose: reading a missing UPPER_CASE key creates an empty
    child node (the reference's add_*_config functions assign into detectron2's default tree, which is not shipped
    here), and merge_from_file accepts new keys (the dumped ckpts/*/config.yaml carry every default explicitly)."""

    def __init__(self, init=None, **_ignored):
        super().__init__()
        object.__setattr__(self, "_frozen", False)
        for k, v in (init or {}).items():
            dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v)

    # attribute access
    def __getattr__(self, k):
        if k.startswith("_"):
            raise AttributeError(k)
        if k in self:
            return self[k]
        if k.isupper() or (k[:1].isupper() and k.replace("_", "").replace("2", "").isalnum() and k.upper() == k):
            if object.__getattribute__(self, "_frozen"):
                raise AttributeError(k)
            child = CfgNode()
            dict.__setitem__(self, k, child)
            return child
        raise AttributeError(k)

    def __setattr__(self, k, v):
        if object.__getattribute__(self, "_frozen"):
            raise AttributeError(f"Attempted to set {k} to {v}, but CfgNode is immutable")
        dict.__setitem__(self, k, CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v)

    def __setitem__(self, k, v):
        self.__setattr__(k, v)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        return out

    def __reduce__(self):
        return (CfgNode, (self.to_dict(),))

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, CfgNode) else v) for k, v in self.items()}

    # yacs API used by the reference
    def freeze(self):
        object.__setattr__(self, "_frozen", True)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.freeze()

    def defrost(self):
        object.__setattr__(self, "_frozen", False)
        for v in self.values():
            if isinstance(v, CfgNode):
                v.defrost()

    def is_frozen(self):
        return object.__getattribute__(self, "_frozen")

    def clone(self):
        return copy.deepcopy(self)

    def dump(self, **kwargs):
        return yaml.safe_dump(self.to_dict(), **kwargs)

    def merge_from_other_cfg(self, other):
        _merge_into(self, other)

    def merge_from_file(self, cfg_filename, allow_unsafe=False):
        _merge_into(self, _load_yaml_with_base(cfg_filename))

    def merge_from_list(self, cfg_list):
        assert len(cfg_list) % 2 == 0, "Override list has odd length"
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                if k not in node:
                    dict.__setitem__(node, k, CfgNode())
                node = node[k]
            if isinstance(v, str):
                try:
                    v = yaml.safe_load(v)
                except yaml.YAMLError:
                    pass
            dict.__setitem__(node, keys[-1], v)


class _LenientLoader(yaml.SafeLoader):
    """The hand-written configs/ YAMLs use `!!python/object/apply:eval [...]`; the value is kept
    as plain data (never evaluated)."""


def _unknown_tag(loader, suffix, node):
    if isinstance(node, yaml.ScalarNode):
        return loader.construct_scalar(node)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_mapping(node, deep=True)


_LenientLoader.add_multi_constructor("tag:yaml.org,2002:python/", _unknown_tag)


def _load_yaml_with_base(path):
    with open(path) as f:
        d = yaml.load(f, Loader=_LenientLoader) or {}
    base = d.pop("_BASE_", None)
    if base:
        bp = base if os.path.isabs(base) else os.path.join(os.path.dirname(path), base)
        merged = _load_yaml_with_base(bp)
        _merge_plain(merged, d)
        return merged
    return d


def _merge_plain(dst, src):
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge_plain(dst[k], v)
        else:
            dst[k] = v


def _merge_into(node, d):
    for k, v in d.items():
        if isinstance(v, dict):
            if not isinstance(node.get(k), CfgNode):
                dict.__setitem__(node, k, CfgNode())
            _merge_into(node[k], v)
        else:
            dict.__setitem__(node, k, v)


def get_cfg():
    """The top-level nodes of detectron2's default tree that the reference touches before merging its YAML
    (detectron2/config/defaults.py; train_net.py:356-360)."""
    cfg = CfgNode()
    cfg.VERSION = 2
    cfg.MODEL.DEVICE = "cuda"
    cfg.MODEL.META_ARCHITECTURE = "GeneralizedRCNN"
    cfg.MODEL.WEIGHTS = ""
    cfg.MODEL.PIXEL_MEAN = [103.530, 116.280, 123.675]
    cfg.MODEL.PIXEL_STD = [1.0, 1.0, 1.0]
    cfg.MODEL.BACKBONE.NAME = "build_resnet_backbone"
    cfg.MODEL.BACKBONE.FREEZE_AT = 2
    cfg.MODEL.SEM_SEG_HEAD.NAME = "SemSegFPNHead"
    cfg.MODEL.SEM_SEG_HEAD.NUM_CLASSES = 54
    cfg.MODEL.SEM_SEG_HEAD.IGNORE_VALUE = 255
    cfg.INPUT.CROP.ENABLED = False
    cfg.INPUT.FORMAT = "BGR"
    cfg.DATASETS.TRAIN = ()
    cfg.DATASETS.TEST = ()
    cfg.DATALOADER.NUM_WORKERS = 4
    cfg.SOLVER.IMS_PER_BATCH = 16
    cfg.TEST.AUG.ENABLED = False
    cfg.OUTPUT_DIR = "./output"
    cfg.SEED = -1
    cfg.CUDNN_BENCHMARK = False
    return cfg


def _called_with_cfg(*args, **kwargs):
    if len(args) and isinstance(args[0], CfgNode):
        return True
    return isinstance(kwargs.get("cfg", None), CfgNode)


def configurable(init_func=None, *, from_config=None):
    """detectron2.config.configurable: `@configurable` on __init__ (class provides from_config) or
    `@configurable(from_config=fn)` on a function."""
    import functools
    if init_func is not None:
        @functools.wraps(init_func)
        def wrapped(self, *args, **kwargs):
            fc = type(self).from_config
            if _called_with_cfg(*args, **kwargs):
                init_func(self, **fc(*args, **kwargs))
            else:
                init_func(self, *args, **kwargs)
        return wrapped

    def wrapper(orig_func):
        @functools.wraps(orig_func)
        def wrapped(*args, **kwargs):
            if _called_with_cfg(*args, **kwargs):
                return orig_func(**from_config(*args, **kwargs))
            return orig_func(*args, **kwargs)
        wrapped.from_config = from_config
        return wrapped
    return wrapper


def add_deeplab_config(cfg):
    """detectron2.projects.deeplab.add_deeplab_config: the keys it adds (train_net.py:358)."""
    cfg.INPUT.CROP.SINGLE_CATEGORY_MAX_AREA = 1.0
    cfg.SOLVER.POLY_LR_POWER = 0.9
    cfg.SOLVER.POLY_LR_CONSTANT_ENDING = 0.0
    cfg.MODEL.SEM_SEG_HEAD.LOSS_TYPE = "hard_pixel_mining"
    cfg.MODEL.SEM_SEG_HEAD.PROJECT_FEATURES = ["res2"]
    cfg.MODEL.SEM_SEG_HEAD.PROJECT_CHANNELS = [48]
    cfg.MODEL.SEM_SEG_HEAD.ASPP_CHANNELS = 256
    cfg.MODEL.SEM_SEG_HEAD.ASPP_DILATIONS = [6, 12, 18]
    cfg.MODEL.SEM_SEG_HEAD.ASPP_DROPOUT = 0.1
    cfg.MODEL.SEM_SEG_HEAD.USE_DEPTHWISE_SEPARABLE_CONV = False
    cfg.MODEL.RESNETS.RES4_DILATION = 1
    cfg.MODEL.RESNETS.RES5_MULTI_GRID = [1, 2, 4]
    cfg.MODEL.RESNETS.STEM_TYPE = "deeplab"


# ----------------------------------------------------------------------------------------------------------------
# registries / modeling (maskformer_model.py:23, swin.py:686, msdeformattn.py:173, train_net.py:74-80)
# ----------------------------------------------------------------------------------------------------------------
class Registry:
    """fvcore.common.registry.Registry"""

    def __init__(self, name):
        self._name = name
        self._obj_map = {}

    def _do_register(self, name, obj):
        self._obj_map[name] = obj          # re-registration overrides (the rba_b200 plug-in point)

    def register(self, obj=None):
        if obj is None:
            def deco(func_or_class):
                self._do_register(func_or_class.__name__, func_or_class)
                return func_or_class
            return deco
        self._do_register(obj.__name__, obj)
        return obj

    def get(self, name):
        ret = self._obj_map.get(name)
        if ret is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return ret

    def __contains__(self, name):
        return name in self._obj_map

    def __iter__(self):
        return iter(self._obj_map.items())


META_ARCH_REGISTRY = Registry("META_ARCH")
BACKBONE_REGISTRY = Registry("BACKBONE")
SEM_SEG_HEADS_REGISTRY = Registry("SEM_SEG_HEADS")
TRANSFORMER_DECODER_REGISTRY = Registry("TRANSFORMER_MODULE")

# name -> class: meta-architectures served by rba_b200 regardless of what the reference registered under that name
PLUGIN_META_ARCH = {}


def build_model(cfg):
    """detectron2.modeling.build_model (train_net.py:78): META_ARCH_REGISTRY lookup + .to(cfg.MODEL.DEVICE).
    The rba_b200 plug-in takes precedence for the architectures it serves."""
    name = cfg.MODEL.META_ARCHITECTURE
    cls = PLUGIN_META_ARCH.get(name) or META_ARCH_REGISTRY.get(name)
    model = cls(cfg)
    model.to(torch.device(cfg.MODEL.DEVICE))
    return model


class ShapeSpec:
    def __init__(self, channels=None, height=None, width=None, stride=None):
        self.channels, self.height, self.width, self.stride = channels, height, width, stride


class Backbone(nn.Module):
    def __init__(self):
        super().__init__()

    @property
    def size_divisibility(self):
        return 0

    def output_shape(self):
        return {name: ShapeSpec(channels=self._out_feature_channels[name], stride=self._out_feature_strides[name])
                for name in self._out_features}


def build_backbone(cfg, input_shape=None):
    if input_shape is None:
        input_shape = ShapeSpec(channels=len(cfg.MODEL.PIXEL_MEAN))
    return BACKBONE_REGISTRY.get(cfg.MODEL.BACKBONE.NAME)(cfg, input_shape)


def build_sem_seg_head(cfg, input_shape):
    return SEM_SEG_HEADS_REGISTRY.get(cfg.MODEL.SEM_SEG_HEAD.NAME)(cfg, input_shape)


def sem_seg_postprocess(result, img_size, output_height, output_width):
    """detectron2.modeling.postprocessing.sem_seg_postprocess (maskformer_model.py:330-333)."""
    result = result[:, : img_size[0], : img_size[1]].expand(1, -1, -1, -1)
    return F.interpolate(result, size=(output_height, output_width), mode="bilinear", align_corners=False)[0]


def get_norm(norm, out_channels):
    if norm is None or norm == "":
        return None
    if isinstance(norm, str):
        norm = {"BN": nn.BatchNorm2d, "SyncBN": nn.BatchNorm2d, "GN": lambda c: nn.GroupNorm(32, c),
                "LN": lambda c: nn.GroupNorm(1, c)}[norm]
    return norm(out_channels)


class Conv2d(nn.Conv2d):
    """detectron2.layers.Conv2d: conv + optional norm + activation."""

    def __init__(self, *args, **kwargs):
        norm = kwargs.pop("norm", None)
        activation = kwargs.pop("activation", None)
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation

    def forward(self, x):
        x = F.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)
        if self.norm is not None:
            x = self.norm(x)
        if self.activation is not None:
            x = self.activation(x)
        return x


class CNNBlockBase(nn.Module):
    def __init__(self, in_channels=0, out_channels=0, stride=1):
        super().__init__()
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride


class ImageList:
    """detectron2.structures.ImageList (maskformer_model.py:257)."""

    def __init__(self, tensor, image_sizes):
        self.tensor = tensor
        self.image_sizes = image_sizes

    def __len__(self):
        return len(self.image_sizes)

    @staticmethod
    def from_tensors(tensors, size_divisibility=0, pad_value=0.0):
        sizes = [tuple(t.shape[-2:]) for t in tensors]
        H, W = max(s[0] for s in sizes), max(s[1] for s in sizes)
        if size_divisibility > 1:
            H = (H + size_divisibility - 1) // size_divisibility * size_divisibility
            W = (W + size_divisibility - 1) // size_divisibility * size_divisibility
        out = tensors[0].new_full((len(tensors),) + tuple(tensors[0].shape[:-2]) + (H, W), pad_value)
        for i, t in enumerate(tensors):
            out[i, ..., : t.shape[-2], : t.shape[-1]].copy_(t)
        return ImageList(out, sizes)


# ----------------------------------------------------------------------------------------------------------------
# detectron2.data catalogs (maskformer_model.py:204; mask2former/data/datasets/register_*.py run at import)
# ----------------------------------------------------------------------------------------------------------------
class _Metadata(types.SimpleNamespace):
    def set(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
        return self

    def get(self, key, default=None):
        return getattr(self, key, default)

    def as_dict(self):
        return dict(self.__dict__)

    def __delattr__(self, k):          # detectron2's builtin datasets are not registered here: deleting is a no-op
        self.__dict__.pop(k, None)

    def __getattr__(self, k):          # unknown metadata reads as None instead of raising at import
        if k.startswith("__"):
            raise AttributeError(k)
        return None


class _MetadataCatalog:
    def __init__(self):
        self._d = {}

    def get(self, name):
        if name not in self._d:
            self._d[name] = _Metadata(name=name)
        return self._d[name]

    def list(self):
        return list(self._d)

    def __contains__(self, name):
        return name in self._d

    def pop(self, name):
        self._d.pop(name, None)

    remove = pop


class _DatasetCatalog:
    def __init__(self):
        self._d = {}

    def register(self, name, func):
        self._d[name] = func

    def get(self, name):
        return self._d[name]()

    def list(self):
        return list(self._d)

    def __contains__(self, name):
        return name in self._d

    def pop(self, name):
        self._d.pop(name, None)

    remove = pop


MetadataCatalog = _MetadataCatalog()
DatasetCatalog = _DatasetCatalog()


# ----------------------------------------------------------------------------------------------------------------
# checkpoint / engine / utils (evaluate_ood.py:108-124, train_net.py:352-366)
# ----------------------------------------------------------------------------------------------------------------
class DetectionCheckpointer:
    """detectron2.checkpoint.DetectionCheckpointer: `.pth` ({"model": state_dict} or a bare state_dict) and the
    `.pkl` model-zoo format ({"model": {name: ndarray}}); strict=False like fvcore's Checkpointer, incompatibilities
    are logged."""

    def __init__(self, model, save_dir="", *, save_to_disk=None, **checkpointables):
        self.model = model
        self.save_dir = save_dir
        self.logger = logging.getLogger("rba_b200.compat.checkpoint")

    def _load_file(self, path):
        if path.endswith(".pkl"):
            with open(path, "rb") as f:
                data = pickle.load(f, encoding="latin1")
            if "model" in data:
                data = data["model"]
            return {"model": {k: torch.as_tensor(np.asarray(v)) for k, v in data.items() if not k.endswith("_momentum")}}
        data = torch.load(path, map_location="cpu", weights_only=False)
        return data if isinstance(data, dict) and "model" in data else {"model": data}

    def load(self, path, checkpointables=None):
        if not path:
            self.logger.info("No checkpoint found. Initializing model from scratch")
            return {}
        if not os.path.isfile(path):
            raise FileNotFoundError(f"Checkpoint {path} not found!")
        ckpt = self._load_file(path)
        sd = {k[len("module."):] if k.startswith("module.") else k: v for k, v in ckpt.pop("model").items()}
        inc = self.model.load_state_dict(sd, strict=False)
        if inc is not None and (inc.missing_keys or inc.unexpected_keys):
            self.logger.warning("checkpoint %s: missing %d keys, unexpected %d keys", path, len(inc.missing_keys),
                                len(inc.unexpected_keys))
        return ckpt

    def resume_or_load(self, path, *, resume=True):
        return self.load(path)

    def save(self, name, **kwargs):
        os.makedirs(self.save_dir or ".", exist_ok=True)
        torch.save({"model": self.model.state_dict(), **kwargs}, os.path.join(self.save_dir or ".", f"{name}.pth"))


def default_setup(cfg, args):
    """detectron2.engine.default_setup (train_net.py:363): output dir + logger; evaluate_ood.py passes an EasyDict
    whose key is 'eval-only' (evaluate_ood.py:112-114)."""
    out = cfg.OUTPUT_DIR if "OUTPUT_DIR" in cfg else None
    if out:
        os.makedirs(out, exist_ok=True)
    setup_logger(out, name="detectron2")


def setup_logger(output=None, distributed_rank=0, *, color=True, name="detectron2", abbrev_name=None, **kwargs):
    logger = logging.getLogger(name)
    logger.setLevel(logging.INFO)
    return logger


class DefaultTrainer:
    """Base of train_net.Trainer (train_net.py:68); only `build_model` is used on the eval path."""

    def __init__(self, cfg=None):
        raise RbaError("training is out of scope of the rba_b200 inference path (detectron2.engine.DefaultTrainer stand-in)")

    @classmethod
    def build_model(cls, cfg):
        return build_model(cfg)

    @classmethod
    def test(cls, cfg, model, evaluators=None):
        raise RbaError("DefaultTrainer.test: detectron2's dataset evaluators are not part of the rba_b200 stand-in")


def default_argument_parser(epilog=None):
    import argparse
    p = argparse.ArgumentParser(epilog=epilog)
    p.add_argument("--config-file", default="", metavar="FILE")
    p.add_argument("--resume", action="store_true")
    p.add_argument("--eval-only", action="store_true")
    p.add_argument("--num-gpus", type=int, default=1)
    p.add_argument("--num-machines", type=int, default=1)
    p.add_argument("--machine-rank", type=int, default=0)
    p.add_argument("--dist-url", default="auto")
    p.add_argument("opts", default=None, nargs=argparse.REMAINDER)
    return p


def launch(main_func, num_gpus_per_machine=1, num_machines=1, machine_rank=0, dist_url=None, args=(), timeout=None):
    return main_func(*args)


class _PathManager:
    """fvcore / detectron2 PathManager: local files only."""

    @staticmethod
    def open(path, mode="r", **kwargs):
        return open(path, mode, **kwargs)

    @staticmethod
    def exists(path):
        return os.path.exists(path)

    @staticmethod
    def isfile(path):
        return os.path.isfile(path)

    @staticmethod
    def isdir(path):
        return os.path.isdir(path)

    @staticmethod
    def ls(path):
        return os.listdir(path)

    @staticmethod
    def mkdirs(path):
        os.makedirs(path, exist_ok=True)

    @staticmethod
    def get_local_path(path, **kwargs):
        return path

    @staticmethod
    def register_handler(handler, allow_override=False):
        pass


def retry_if_cuda_oom(func):
    return func


# ----------------------------------------------------------------------------------------------------------------
# small third parties
# ----------------------------------------------------------------------------------------------------------------
class EasyDict(dict):
    """easydict.EasyDict (evaluate_ood.py:23,112; support.py)."""

    def __init__(self, d=None, **kwargs):
        super().__init__()
        d = dict(d or {})
        d.update(kwargs)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, name, value):
        if isinstance(value, (list, tuple)):
            value = type(value)(self.__class__(x) if isinstance(x, dict) else x for x in value)
        elif isinstance(value, dict) and not isinstance(value, EasyDict):
            value = EasyDict(value)
        super().__setattr__(name, value)
        super().__setitem__(name, value)

    __setitem__ = __setattr__

    def update(self, e=None, **f):
        d = dict(e or {})
        d.update(f)
        for k in d:
            setattr(self, k, d[k])

    def pop(self, k, *args):
        if hasattr(self, k):
            delattr(self, k)
        return super().pop(k, *args)


class _ACompose:
    """albumentations.Compose (support.py:73-81): transforms applied in order to `image` (HWC ndarray) and `mask`."""

    def __init__(self, transforms, **kwargs):
        self.transforms = list(transforms)

    def __call__(self, force_apply=False, **data):
        for t in self.transforms:
            data = t(**data)
        return data


class _AResize:
    """albumentations.Resize: bilinear for the image, nearest for the mask (cv2, like albumentations)."""

    def __init__(self, height, width, interpolation=1, always_apply=False, p=1):
        self.height, self.width, self.interpolation = height, width, interpolation

    def __call__(self, **data):
        import cv2
        if data.get("image") is not None:
            data["image"] = cv2.resize(data["image"], (self.width, self.height), interpolation=self.interpolation)
        if data.get("mask") is not None:
            data["mask"] = cv2.resize(data["mask"], (self.width, self.height), interpolation=cv2.INTER_NEAREST)
        return data


class _AToTensorV2:
    """albumentations.pytorch.ToTensorV2: HWC ndarray -> CHW tensor (dtype kept), mask -> tensor."""

    def __init__(self, transpose_mask=False, always_apply=True, p=1.0):
        self.transpose_mask = transpose_mask

    def __call__(self, **data):
        img = data.get("image")
        if img is not None:
            if img.ndim == 2:
                img = img[:, :, None]
            data["image"] = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1)))
        m = data.get("mask")
        if m is not None:
            if self.transpose_mask and m.ndim == 3:
                m = m.transpose(2, 0, 1)
            data["mask"] = torch.from_numpy(np.ascontiguousarray(m))
        return data


def fpr_at_95_tpr(preds, labels, pos_label=1):
    """ood_metrics.fpr_at_95_tpr (imported by support.py:24): FPR at the first threshold whose TPR >= 0.95."""
    from sklearn.metrics import roc_curve
    fpr, tpr, _ = roc_curve(labels, preds, pos_label=pos_label)
    if all(tpr < 0.95):
        return 0.0
    if all(tpr >= 0.95):
        return float(fpr[np.argmin(fpr)])
    return float(np.interp(0.95, tpr, fpr))


def _imsave(fname, arr, cmap=None, **kwargs):
    """matplotlib.image.imsave for --store_anomaly_scores (evaluate_ood.py:225): min-max normalised grey PNG."""
    from PIL import Image
    a = np.asarray(arr, dtype=np.float64)
    lo, hi = float(a.min()), float(a.max())
    a = (a - lo) / (hi - lo) if hi > lo else np.zeros_like(a)
    Image.fromarray((a * 255).astype(np.uint8)).save(fname)


def _trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class _DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return x          # inference: identity


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def _c2_xavier_fill(module):
    nn.init.kaiming_uniform_(module.weight, a=1)
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


def _c2_msra_fill(module):
    nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
    if module.bias is not None:
        nn.init.constant_(module.bias, 0)


# ----------------------------------------------------------------------------------------------------------------
# module table + importer
# ----------------------------------------------------------------------------------------------------------------
def _functional_modules():
    """dotted module name -> {attribute: object} for everything with real behaviour; all other attributes of modules
    under STANDIN_ROOTS are inert placeholders."""
    comm = dict(get_world_size=lambda: 1, get_rank=lambda: 0, get_local_rank=lambda: 0, is_main_process=lambda: True,
                synchronize=lambda: None, all_gather=lambda data, group=None: [data], gather=lambda data, dst=0, group=None: [data],
                shared_random_seed=lambda: 0, reduce_dict=lambda d, average=True: d)
    return {
        "detectron2": {"__version__": "0.6+rba_b200.compat"},
        "detectron2.config": dict(CfgNode=CfgNode, get_cfg=get_cfg, configurable=configurable),
        "detectron2.modeling": dict(META_ARCH_REGISTRY=META_ARCH_REGISTRY, BACKBONE_REGISTRY=BACKBONE_REGISTRY,
                                    SEM_SEG_HEADS_REGISTRY=SEM_SEG_HEADS_REGISTRY, Backbone=Backbone, ShapeSpec=ShapeSpec,
                                    build_model=build_model, build_backbone=build_backbone,
                                    build_sem_seg_head=build_sem_seg_head),
        "detectron2.modeling.postprocessing": dict(sem_seg_postprocess=sem_seg_postprocess),
        "detectron2.modeling.backbone": dict(Backbone=Backbone, BACKBONE_REGISTRY=BACKBONE_REGISTRY),
        "detectron2.layers": dict(Conv2d=Conv2d, ShapeSpec=ShapeSpec, get_norm=get_norm, CNNBlockBase=CNNBlockBase),
        "detectron2.structures": dict(ImageList=ImageList),
        "detectron2.data": dict(MetadataCatalog=MetadataCatalog, DatasetCatalog=DatasetCatalog),
        "detectron2.data.catalog": dict(MetadataCatalog=MetadataCatalog, DatasetCatalog=DatasetCatalog),
        "detectron2.checkpoint": dict(DetectionCheckpointer=DetectionCheckpointer),
        "detectron2.engine": dict(DefaultTrainer=DefaultTrainer, default_argument_parser=default_argument_parser,
                                  default_setup=default_setup, launch=launch),
        "detectron2.utils.comm": comm,
        "detectron2.utils.registry": dict(Registry=Registry),
        "detectron2.utils.logger": dict(setup_logger=setup_logger, create_small_table=lambda d: str(d)),
        "detectron2.utils.memory": dict(retry_if_cuda_oom=retry_if_cuda_oom),
        "detectron2.utils.file_io": dict(PathManager=_PathManager),
        "detectron2.projects.deeplab": dict(add_deeplab_config=add_deeplab_config),
        "fvcore.common.registry": dict(Registry=Registry),
        "fvcore.common.file_io": dict(PathManager=_PathManager),
        "fvcore.nn.weight_init": dict(c2_xavier_fill=_c2_xavier_fill, c2_msra_fill=_c2_msra_fill),
        "timm.models.layers": dict(DropPath=_DropPath, to_2tuple=_to_2tuple, trunc_normal_=_trunc_normal_),
        "timm.models.registry": dict(register_model=lambda fn: fn),
        "timm.models.vision_transformer": dict(_cfg=lambda url="", **kw: dict(url=url, **kw)),
        "easydict": dict(EasyDict=EasyDict),
        "albumentations": dict(Compose=_ACompose, Resize=_AResize),
        "albumentations.pytorch": dict(ToTensorV2=_AToTensorV2),
        "matplotlib.image": dict(imsave=_imsave),
        "ood_metrics": dict(fpr_at_95_tpr=fpr_at_95_tpr),
    }


class _StandinFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self, roots):
        self.roots = set(roots)
        self.table = _functional_modules()

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] not in self.roots:
            return None
        return importlib.machinery.ModuleSpec(fullname, self, is_package=True)

    def create_module(self, spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        m.__rba_b200_standin__ = True
        return m

    def exec_module(self, module):
        for k, v in self.table.get(module.__name__, {}).items():
            setattr(module, k, v)


_finder = None


def _really_importable(root):
    for f in sys.meta_path:
        if f is _finder:
            continue
        try:
            if f.find_spec(root, None) is not None:
                return True
        except Exception:
            pass
    return False


def install(roots=STANDIN_ROOTS):
    """Makes the stand-ins importable for every root package that is not really installed.  Idempotent.
    Returns the list of roots that are served by stand-ins."""
    global _finder
    if _finder is None:
        served = [r for r in roots if r not in sys.modules and not _really_importable(r)]
        _finder = _StandinFinder(served)
        sys.meta_path.append(_finder)       # after the real finders: a real install always wins
    return sorted(_finder.roots)
