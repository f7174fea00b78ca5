"""ResNet backbone (BASELINE.json configs[0]: "ResNet-50 Mask2Former 1dl, 1x512x1024 synthetic, CPU-only forward + RbA
score").  detectron2's build_resnet_backbone is neither vendored by the reference nor version-pinned: PARITY
UNPINNED against detectron2 itself.  What IS pinned here:
  * the oracle's restatement (oracle/rba_oracle.py::resnet_forward) and the stand-in module the reference is built with
    (oracle/ref_shims/detectron2/modeling/backbone/resnet.py) against torchvision's resnet50 / resnet101 under the key
    mapping of the reference's own tools/convert-torchvision-to-d2.py:33-44;
  * the whole oracle forward against the LIVE reference (its own MaskFormer / pixel decoder / transformer decoder modules
    around that stand-in backbone), when the reference checkout is present."""
import re

import pytest
import torch
import torchvision

import ref_loader
import rba_oracle as O
from rba_b200 import config, weights


def tv_to_d2(k):
    """tools/convert-torchvision-to-d2.py:33-44, plus the `backbone.` prefix MaskFormer gives its backbone."""
    if "layer" not in k:
        k = "stem." + k
    for t in [1, 2, 3, 4]:
        k = k.replace(f"layer{t}", f"res{t + 1}")
    for t in [1, 2, 3]:
        k = k.replace(f"bn{t}", f"conv{t}.norm")
    k = k.replace("downsample.0", "shortcut").replace("downsample.1", "shortcut.norm")
    return "backbone." + k


@pytest.mark.parametrize("depth", [50, 101])
def test_resnet_oracle_matches_torchvision(depth):
    torch.manual_seed(depth)
    tv = getattr(torchvision.models, f"resnet{depth}")(weights=None).eval()
    with torch.no_grad():                                   # non-trivial batch-norm statistics and affines
        for m in tv.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.normal_(1, 0.1)
                m.bias.normal_(0, 0.1)
    sd = {tv_to_d2(k): v for k, v in tv.state_dict().items() if not k.startswith("fc.")}
    mc = config.r50_1dl(depth)
    want = {k for k in weights.param_specs(mc) if k.startswith("backbone.")}
    assert set(sd) == want, (sorted(set(sd) ^ want)[:6])   # the key inventory IS the converter's output
    x = torch.randn(2, 3, 96, 160)
    with torch.no_grad():
        y = tv.maxpool(tv.relu(tv.bn1(tv.conv1(x))))
        ref = {}
        for i, layer in enumerate([tv.layer1, tv.layer2, tv.layer3, tv.layer4]):
            y = layer(y)
            ref[f"res{i + 2}"] = y
    got = O.resnet_forward(sd, mc, x)
    for k in ref:
        assert got[k].shape == ref[k].shape
        assert (got[k] - ref[k]).abs().max() <= 1e-5 * max(1.0, float(ref[k].abs().max())), k


@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_oracle_r50_matches_live_reference_with_standin_backbone():
    """MaskFormer(cfg) of the reference with MODEL.BACKBONE.NAME build_resnet_backbone (stand-in), 1dl, vs the oracle."""
    mc = config.r50_1dl()
    cfg = ref_loader.load_cfg("swin_b_1dl", ref_loader.r50_overrides(dec_layers=mc.dec_layers, levels=1))
    model = ref_loader.build_reference_model(cfg, seed=0)
    sd = weights.init_state_dict(mc, seed=29, perturb=0.02)
    ref_loader.load_state_dict_into(model, sd)
    g = torch.Generator().manual_seed(7)
    img = torch.randint(0, 256, (3, 96, 160), dtype=torch.uint8, generator=g)
    with torch.no_grad():
        sem = model([{"image": img}])[0]["sem_seg"]
    out = O.forward(sd, mc, [img])
    assert (out["sem_seg"][0] - sem).abs().max() < 2e-5
    assert (out["rba"][0] - (-sem.tanh().sum(0))).abs().max() < 2e-5


def test_config_accepts_resnet_yaml_keys():
    y = {"MODEL": {"BACKBONE": {"NAME": "build_resnet_backbone"}, "RESNETS": {"DEPTH": 50, "STRIDE_IN_1X1": False},
                   "PIXEL_MEAN": [123.675, 116.28, 103.53], "PIXEL_STD": [58.395, 57.12, 57.375],
                   "SEM_SEG_HEAD": {"PIXEL_DECODER_NAME": "MSDeformAttnPixelDecoder", "NORM": "GN", "CONVS_DIM": 256, "MASK_DIM": 256,
                                    "NUM_CLASSES": 19, "IN_FEATURES": ["res2", "res3", "res4", "res5"],
                                    "DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES": ["res5"], "COMMON_STRIDE": 4,
                                    "TRANSFORMER_ENC_LAYERS": 6},
                   "MASK_FORMER": {"TRANSFORMER_DECODER_NAME": "MultiScaleMaskedTransformerDecoder", "PRE_NORM": False, "NHEADS": 8,
                                   "HIDDEN_DIM": 256, "DIM_FEEDFORWARD": 2048, "DEC_LAYERS": 2, "NUM_OBJECT_QUERIES": 100,
                                   "SIZE_DIVISIBILITY": 32}}}
    mc = config.model_config_from_cfg(y)
    assert mc.backbone == "resnet" and mc.feature_channels == [256, 512, 1024, 2048] and mc.dec_layers == 1
    y["MODEL"]["RESNETS"]["STRIDE_IN_1X1"] = True
    with pytest.raises(ValueError):
        config.model_config_from_cfg(y)
    assert re.match(r"backbone\.res2\.0\.shortcut\.weight", [k for k in weights.param_specs(mc) if "shortcut" in k][0])
