"""Model configuration: the subset of the reference's yacs keys that the hot path reads
(mask2former/config.py:40-90,170-172) and their mapping to the C-ABI `rba_config`."""
from dataclasses import dataclass, field
from typing import List

import yaml

from ._lib import RbaConfig


@dataclass
class ModelConfig:
    embed_dim: int = 128
    depths: List[int] = field(default_factory=lambda: [2, 2, 18, 2])
    num_heads: List[int] = field(default_factory=lambda: [4, 8, 16, 32])
    window_size: int = 12
    mlp_ratio: float = 4.0
    patch_size: int = 4
    conv_dim: int = 256
    mask_dim: int = 256
    num_classes: int = 19
    in_features: List[str] = field(default_factory=lambda: ["res2", "res3", "res4", "res5"])
    transformer_in_features: List[str] = field(default_factory=lambda: ["res5"])
    common_stride: int = 4
    enc_layers: int = 6
    enc_heads: int = 8
    enc_points: int = 4
    enc_ffn: int = 1024          # hard-coded in the reference, msdeformattn.py:315
    hidden_dim: int = 256
    nheads: int = 8
    dim_feedforward: int = 2048
    dec_layers: int = 1          # MODEL.MASK_FORMER.DEC_LAYERS - 1 (mask2former_transformer_decoder.py:387-388)
    num_queries: int = 100
    size_divisibility: int = 32
    pixel_mean: List[float] = field(default_factory=lambda: [123.675, 116.28, 103.53])
    pixel_std: List[float] = field(default_factory=lambda: [58.395, 57.12, 57.375])
    backbone: str = "swin"         # "swin" (D2SwinTransformer) or "resnet" (detectron2 build_resnet_backbone, bottleneck R50 / R101)
    resnet_depth: int = 50         # MODEL.RESNETS.DEPTH
    ood_prediction: bool = False   # MODEL.MASK_FORMER.DENSE_HYBRID_LOSS: predictor.ood_pred head (mask2former_transformer_decoder.py:365-366,394)

    @property
    def num_enc_levels(self):
        return len(self.transformer_in_features)

    @property
    def feature_channels(self):
        """Channels of res2..res5 as the pixel decoder sees them."""
        if self.backbone == "resnet":
            return [256, 512, 1024, 2048]
        return [self.embed_dim << i for i in range(4)]

    def validate(self):
        if self.backbone not in ("swin", "resnet"):
            raise ValueError(f"unknown backbone {self.backbone!r}")
        if self.backbone == "resnet" and self.resnet_depth not in (50, 101):
            raise ValueError("only bottleneck ResNet-50 / ResNet-101 are built")
        if self.mlp_ratio != 4.0 or self.patch_size != 4:
            raise ValueError("only MLP_RATIO 4.0 / PATCH_SIZE 4 are supported")
        if self.hidden_dim != self.conv_dim:
            raise ValueError("HIDDEN_DIM must equal CONVS_DIM (decoder input_proj is not built)")
        if sorted(self.transformer_in_features) not in (["res5"], ["res3", "res4", "res5"]):
            raise ValueError(f"unsupported DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES {self.transformer_in_features}")
        if sorted(self.in_features) != ["res2", "res3", "res4", "res5"] or self.common_stride != 4:
            raise ValueError("IN_FEATURES res2..res5 with COMMON_STRIDE 4 expected")
        return self

    def to_ctypes(self):
        c = RbaConfig()
        c.embed_dim = self.embed_dim
        for i in range(4):
            c.depths[i] = self.depths[i]
            c.num_heads[i] = self.num_heads[i]
        c.window_size = self.window_size
        c.conv_dim, c.mask_dim, c.num_classes = self.conv_dim, self.mask_dim, self.num_classes
        c.num_queries, c.nheads, c.dim_feedforward = self.num_queries, self.nheads, self.dim_feedforward
        c.dec_layers, c.enc_layers, c.enc_points, c.enc_ffn = self.dec_layers, self.enc_layers, self.enc_points, self.enc_ffn
        c.num_enc_levels = self.num_enc_levels
        c.size_divisibility = self.size_divisibility
        c.backbone_type = 1 if self.backbone == "resnet" else 0
        c.resnet_depth = self.resnet_depth
        for i in range(3):
            c.pixel_mean[i] = self.pixel_mean[i]
            c.pixel_std[i] = self.pixel_std[i]
        return c


def _get(node, dotted):
    for k in dotted.split("."):
        node = node[k] if isinstance(node, dict) else getattr(node, k)
    return node


def _get_default(node, dotted, default):
    try:
        return _get(node, dotted)
    except (KeyError, AttributeError):
        return default


def model_config_from_cfg(cfg):
    """cfg: a yacs-like CfgNode / nested dict with the reference's key names (e.g. yaml.safe_load of
    ckpts/<name>/config.yaml).  Raises on architectures outside the built hot path."""
    M = cfg["MODEL"] if isinstance(cfg, dict) else cfg.MODEL
    g = lambda k: _get(M, k)  # noqa: E731
    if g("BACKBONE.NAME") not in ("D2SwinTransformer", "build_resnet_backbone"):
        raise ValueError(f"backbone {g('BACKBONE.NAME')} is not built (D2SwinTransformer and build_resnet_backbone are)")
    resnet = g("BACKBONE.NAME") == "build_resnet_backbone"
    if resnet:
        if _get_default(M, "RESNETS.STRIDE_IN_1X1", False) or _get_default(M, "RESNETS.NUM_GROUPS", 1) != 1 or \
                any(_get_default(M, "RESNETS.DEFORM_ON_PER_STAGE", [False])):
            raise ValueError("only plain bottleneck ResNets with STRIDE_IN_1X1: False are built")
    if g("SEM_SEG_HEAD.PIXEL_DECODER_NAME") != "MSDeformAttnPixelDecoder":
        raise ValueError("only MSDeformAttnPixelDecoder is built")
    if g("MASK_FORMER.TRANSFORMER_DECODER_NAME") != "MultiScaleMaskedTransformerDecoder":
        raise ValueError("only MultiScaleMaskedTransformerDecoder is built")
    if g("MASK_FORMER.PRE_NORM") or (not resnet and (g("SWIN.APE") or not g("SWIN.QKV_BIAS") or not g("SWIN.PATCH_NORM"))):
        raise ValueError("PRE_NORM / APE / no QKV_BIAS / no PATCH_NORM variants are not built")
    if g("SEM_SEG_HEAD.NORM") != "GN":
        raise ValueError("SEM_SEG_HEAD.NORM must be GN")
    sw = {} if resnet else dict(
        embed_dim=g("SWIN.EMBED_DIM"), depths=list(g("SWIN.DEPTHS")), num_heads=list(g("SWIN.NUM_HEADS")),
        window_size=g("SWIN.WINDOW_SIZE"), mlp_ratio=float(g("SWIN.MLP_RATIO")), patch_size=g("SWIN.PATCH_SIZE"))
    mc = ModelConfig(
        backbone="resnet" if resnet else "swin", resnet_depth=int(_get_default(M, "RESNETS.DEPTH", 50)), **sw,
        conv_dim=g("SEM_SEG_HEAD.CONVS_DIM"), mask_dim=g("SEM_SEG_HEAD.MASK_DIM"), num_classes=g("SEM_SEG_HEAD.NUM_CLASSES"),
        in_features=list(g("SEM_SEG_HEAD.IN_FEATURES")),
        transformer_in_features=list(g("SEM_SEG_HEAD.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES")),
        common_stride=g("SEM_SEG_HEAD.COMMON_STRIDE"), enc_layers=g("SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS"),
        enc_heads=g("MASK_FORMER.NHEADS"), hidden_dim=g("MASK_FORMER.HIDDEN_DIM"), nheads=g("MASK_FORMER.NHEADS"),
        dim_feedforward=g("MASK_FORMER.DIM_FEEDFORWARD"), dec_layers=g("MASK_FORMER.DEC_LAYERS") - 1,
        num_queries=g("MASK_FORMER.NUM_OBJECT_QUERIES"), size_divisibility=g("MASK_FORMER.SIZE_DIVISIBILITY"),
        pixel_mean=list(g("PIXEL_MEAN")), pixel_std=list(g("PIXEL_STD")),
        ood_prediction=bool(_get_default(M, "MASK_FORMER.DENSE_HYBRID_LOSS", False)),
    )
    return mc.validate()


def model_config_from_yaml(path):
    with open(path) as f:
        return model_config_from_cfg(yaml.safe_load(f))


# The shipped checkpoint architectures (ckpts/swin_b_1dl/config.yaml, ckpts/swin_l_1dl/config.yaml) as presets,
# so they can be built where the reference checkout (and its YAML files) is not present.
def swin_b_1dl():
    return ModelConfig().validate()


def swin_l_1dl():
    return ModelConfig(embed_dim=192, num_heads=[6, 12, 24, 48]).validate()


def swin_b_full(dec_layers=9):
    """3-level / full-decoder variant (configs/.../maskformer2_R50_bs16_90k.yaml:15,35 inherited by the swin_base yaml)."""
    return ModelConfig(transformer_in_features=["res3", "res4", "res5"], dec_layers=dec_layers).validate()


def r50_1dl(depth=50):
    """ResNet-50 single-decoder-layer variant (BASELINE.json configs[0]): maskformer2_R50_bs16_90k.yaml with the two overrides
    of maskformer2_R101_bs16_90k_1dl.yaml:13-15 (DEC_LAYERS 2, encoder on res5 only); depth=101 is that R101 file itself."""
    return ModelConfig(backbone="resnet", resnet_depth=depth).validate()


def r50_full(dec_layers=9, depth=50):
    """maskformer2_R50_bs16_90k.yaml as is: 3 encoder levels, DEC_LAYERS 10."""
    return ModelConfig(backbone="resnet", resnet_depth=depth, transformer_in_features=["res3", "res4", "res5"],
                       dec_layers=dec_layers).validate()


def tiny_test(depths=(2, 2, 2, 2), levels=1, dec_layers=1, ood_prediction=False):
    """Small Swin (embed 32) for fast parity tests; same code paths as Swin-B."""
    tin = ["res5"] if levels == 1 else ["res3", "res4", "res5"]
    return ModelConfig(embed_dim=32, depths=list(depths), num_heads=[1, 2, 4, 8], transformer_in_features=tin,
                       dec_layers=dec_layers, enc_layers=2, ood_prediction=ood_prediction).validate()
