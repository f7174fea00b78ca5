"""CPU: the C-ABI library builds, loads and exports every symbol include/rba_b200.h declares; the product path
refuses to run without a GPU (no CPU fallback); host-side logic (config, weights inventory, state_dict handling)."""
import ctypes
import os
import re

import pytest
import torch

import rba_b200
from rba_b200 import _lib, config, weights

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "rba_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rba_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _header_functions()
    assert len(names) >= 20
    L = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(L, n), f"{n} declared in rba_b200.h but not exported by librba_b200.so"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_lib.PROTOTYPES) == names
    assert _lib.lib().rba_version() >= 1
    assert _lib.launch_count() >= 0


def test_struct_layouts_match_header():
    # rba_config: 22 int32 + 6 float + backbone_type, resnet_depth; rba_gemm_args must round-trip through the C side unchanged
    # (checked on GPU by use)
    assert ctypes.sizeof(_lib.RbaConfig) == 4 * (1 + 4 + 4 + 1 + 12) + 4 * 6 + 4 * 2
    assert ctypes.sizeof(_lib.RbaGemmArgs) % 8 == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_product_fails_loudly_without_gpu():
    mc = config.tiny_test()
    with pytest.raises(rba_b200.RbaError):
        rba_b200.Engine(mc)
    m = rba_b200.MaskFormer(mc)
    with pytest.raises(rba_b200.RbaError):
        m([{"image": torch.zeros(3, 32, 32, dtype=torch.uint8)}])
    with pytest.raises(rba_b200.RbaError):
        rba_b200.ops.score_fused(torch.zeros(1, 4, 2, 2), torch.zeros(1, 4, 20), (8, 8))
    # the C entry point itself reports the missing device instead of computing anything
    h = ctypes.c_void_p()
    cfg = mc.to_ctypes()
    rc = _lib.lib().rba_model_create(ctypes.byref(cfg), 0, ctypes.byref(h))
    assert rc != 0 and b"no CUDA device" in _lib.lib().rba_last_error()


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, "rba_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dp, f)).read()
                assert "rba_oracle" not in s and "ref_loader" not in s and "/root/reference" not in s, f


def test_config_presets_and_validation():
    b, l = config.swin_b_1dl(), config.swin_l_1dl()
    assert b.embed_dim == 128 and l.embed_dim == 192 and b.num_enc_levels == 1 and b.dec_layers == 1
    assert config.swin_b_full().num_enc_levels == 3
    y = {"MODEL": {"BACKBONE": {"NAME": "D2MixVisionTransformer"}}}
    with pytest.raises(ValueError):
        config.model_config_from_cfg(y)
    r = config.r50_1dl()
    assert r.backbone == "resnet" and r.to_ctypes().backbone_type == 1 and r.to_ctypes().resnet_depth == 50
    assert config.r50_full().num_enc_levels == 3 and config.r50_1dl(101).resnet_depth == 101
    c = b.to_ctypes()
    assert c.backbone_type == 0 and list(c.depths) == [2, 2, 18, 2] and c.num_enc_levels == 1 and abs(c.pixel_std[2] - 57.375) < 1e-6


def test_weight_inventory_and_state_dict_roundtrip():
    mc = config.tiny_test(levels=3, dec_layers=2)
    sd = weights.init_state_dict(mc, seed=3, perturb=0.01)
    specs = weights.param_specs(mc)
    assert list(sd) == list(specs)
    for k, (shape, _) in specs.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    m = rba_b200.MaskFormer(mc)
    res = m.load_state_dict(sd)
    assert not res.missing_keys and not res.unexpected_keys
    got = m.state_dict()
    assert all(torch.equal(got[k], sd[k]) for k in sd)
    # legacy checkpoints (mask_former_head.py:31-53, mask2former_transformer_decoder.py:237-258)
    legacy = {}
    for k, v in sd.items():
        k2 = k.replace("sem_seg_head.pixel_decoder.", "sem_seg_head.").replace("query_feat", "static_query")
        legacy[k2] = v
    m2 = rba_b200.MaskFormer(mc)
    m2.load_state_dict(legacy)
    assert all(torch.equal(m2.state_dict()[k], sd[k]) for k in sd)
    bad = dict(sd)
    bad["backbone.norm0.weight"] = torch.zeros(7)
    with pytest.raises(RuntimeError, match="size mismatch"):
        m.load_state_dict(bad)
    with pytest.raises(RuntimeError):
        m.load_state_dict({k: v for k, v in sd.items() if "norm0" not in k})
    m.load_state_dict({k: v for k, v in sd.items() if "norm0" not in k}, strict=False)
    assert weights.relative_position_index(12).shape == (144, 144)
