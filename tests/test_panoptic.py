"""Open-panoptic inference: rba_b200.panoptic.panoptic_inference against the reference's own
MaskFormer.panoptic_inference (maskformer_model.py:394-481; tests/golden/panoptic.pt, oracle/make_golden_panoptic.py).  Host logic:
runs on CPU tensors here and on CUDA tensors on the GPU box."""
import pytest
import torch

from conftest import load_golden
from make_golden_panoptic import make_inputs
from rba_b200.panoptic import panoptic_inference


def _run(name, f, things, device):
    c = f["case"]
    cls, masks = make_inputs(c)
    assert abs(float(cls.double().sum() + masks.double().sum()) - f["in_checksum"]) < 1e-6 * max(1.0, abs(f["in_checksum"]))
    out = panoptic_inference(cls.to(device), masks.to(device), c["K"], 0.8, 0.8, things, open_panoptic=c["open"],
                             ood_threshold=c["thr"], pixel_min=c["pmin"], return_ood_pred=(name == "open_ret"))
    assert torch.equal(out[0].cpu(), f["panoptic_seg"]), name
    assert out[1] == f["segments_info"], name
    if name == "open_ret":
        assert (out[2].cpu() - f["ood_mask"]).abs().max() < 1e-5
    return out


@pytest.mark.parametrize("name", ["closed", "open", "open_ret", "nothing_kept"])
def test_panoptic_inference_matches_reference(name):
    fix = load_golden("panoptic.pt")
    _run(name, fix["cases"][name], fix["things"], "cpu")


def test_panoptic_cases_are_not_trivial():
    fix = load_golden("panoptic.pt")["cases"]
    assert len(fix["closed"]["segments_info"]) >= 3 and len(fix["open"]["segments_info"]) >= 3
    assert any(s["category_id"] == 255 for s in fix["open"]["segments_info"])          # an OoD segment was created
    assert fix["nothing_kept"]["segments_info"] == []


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["closed", "open"])
def test_panoptic_inference_on_gpu(dev, name):
    fix = load_golden("panoptic.pt")
    _run(name, fix["cases"][name], fix["things"], dev)
