"""The reference's evaluate_ood.py, UNMODIFIED, end to end through rba_b200.compat (SURVEY §8b "compat surface"):
config -> train_net.setup -> Trainer.build_model (-> rba_b200.MaskFormer) -> DetectionCheckpointer -> the reference's
datasets / DataLoader / OODEvaluator -> results.pkl, on a synthetic dataset tree.  Needs the reference checkout
(/root/reference: build container only); the forward is the oracle's on CPU and the real engine on a GPU."""
import json
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch
import yaml

from rba_b200 import compat, weights
from rba_b200.compat import standins

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the reference checkout (build container) or its verbatim copy shipped to the GPU box (tools/make_baseline_ref.py)
REF = os.environ.get("RBA_REFERENCE_ROOT") or next(
    (p for p in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")) if os.path.isfile(os.path.join(p, "evaluate_ood.py"))),
    "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "evaluate_ood.py")), reason="reference checkout not present")


def _write_png(path, arr):
    import cv2
    os.makedirs(os.path.dirname(path), exist_ok=True)
    assert cv2.imwrite(path, arr)


def make_dataset_tree(root, h=64, w=96, seed=0, n_laf=2, ra_hw=None):
    """Smallest tree the reference's get_datasets() accepts (support.py:26-93): two RoadAnomaly frames, two
    FishyscapesLAF frames, empty folders for the seven datasets the default run constructs but does not evaluate."""
    rng = np.random.default_rng(seed)
    for d in ["cityscapes/leftImg8bit/val", "cityscapes/gtFine/val", "bdd100k/seg", "Fishyscapes/fs_val_v1",
              "Fishyscapes/fs_static_images_v1", "Fishyscapes/fs_val_v2", "Fishyscapes/fs_static_images_v2",
              "SegmentMeIfYouCan/dataset_AnomalyTrack/images", "SegmentMeIfYouCan/dataset_AnomalyTrack/labels_masks",
              "SegmentMeIfYouCan/dataset_ObstacleTrack/images", "SegmentMeIfYouCan/dataset_ObstacleTrack/labels_masks",
              "LostAndFound/leftImg8bit/test"]:
        os.makedirs(os.path.join(root, d), exist_ok=True)
    open(os.path.join(root, "bdd100k/seg/val_paths.txt"), "w").close()
    ra = os.path.join(root, "RoadAnomaly/RoadAnomaly_jpg")
    names = ["frame0.jpg", "frame1.jpg"]
    os.makedirs(ra, exist_ok=True)
    with open(os.path.join(ra, "frame_list.json"), "w") as f:
        json.dump(names, f)
    rh, rw = ra_hw or (h, w)
    for n in names:
        _write_png(os.path.join(ra, "frames", n), rng.integers(0, 256, (rh, rw, 3), dtype=np.uint8))
        lab = np.zeros((rh, rw, 3), np.uint8)
        lab[rh // 4: rh // 2, rw // 4: rw // 2] = 2        # anomaly (2 -> 1 in road_anomaly.py:41)
        _write_png(os.path.join(ra, "frames", n[:-4] + ".labels", "labels_semantic.png"), lab)
    fs = os.path.join(root, "Fishyscapes")
    for i in range(n_laf):
        stem = f"city_{i:06d}_000019_"
        _write_png(os.path.join(fs, "laf_images", stem + "leftImg8bit.png"), rng.integers(0, 256, (h, w, 3), dtype=np.uint8))
        lab = np.zeros((h, w, 3), np.uint8)
        lab[:8] = 255                                      # ignored region
        lab[h // 2:, w // 2:] = 1                          # anomaly
        _write_png(os.path.join(fs, "fishyscapes_lostandfound", f"{i:04d}_{stem}labels.png"), lab)


def make_models_folder(root, device, arch="tiny"):
    """<root>/<arch>/{config.yaml, model_final.pth}: the shipped swin_b_1dl config (every detectron2 default spelled out),
    for arch "tiny" shrunk to the tiny test architecture, for "swin_b_full" switched to the 3-level / 9-layer decoder
    (maskformer2_R50_bs16_90k.yaml:15,35), with random-init weights in the reference's state_dict layout."""
    with open(os.path.join(REF, "ckpts", "swin_b_1dl", "config.yaml")) as f:
        y = yaml.safe_load(f)
    if arch == "tiny":
        y["MODEL"]["SWIN"].update(EMBED_DIM=32, DEPTHS=[2, 2, 2, 2], NUM_HEADS=[1, 2, 4, 8])
        y["MODEL"]["SEM_SEG_HEAD"]["TRANSFORMER_ENC_LAYERS"] = 2
    elif arch == "swin_b_full":
        y["MODEL"]["SEM_SEG_HEAD"]["DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES"] = ["res3", "res4", "res5"]
        y["MODEL"]["MASK_FORMER"]["DEC_LAYERS"] = 10
    else:
        assert arch == "swin_b_1dl", arch
    y["MODEL"]["DEVICE"] = device
    y["MODEL"]["WEIGHTS"] = ""
    d = os.path.join(root, arch)
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "config.yaml"), "w") as f:
        yaml.safe_dump(y, f)
    from rba_b200.config import model_config_from_cfg
    mc = model_config_from_cfg(y)
    sd = weights.init_state_dict(mc, seed=5, perturb=0.02)
    torch.save({"model": sd}, os.path.join(d, "model_final.pth"))
    return mc, sd


@needs_ref
def test_evaluate_ood_runs_unchanged(tmp_path):
    data, models, out = str(tmp_path / "data"), str(tmp_path / "models"), str(tmp_path / "results")
    make_dataset_tree(data)
    device = "cuda" if torch.cuda.is_available() else "cpu"
    make_models_folder(models, device)
    cmd = [sys.executable, os.path.join(ROOT, "tests", "helpers", "run_reference_script.py"), REF, "evaluate_ood.py",
           "--datasets_folder", data, "--models_folder", models, "--out_path", out, "--num_workers", "0",
           "--device", device]
    env = dict(os.environ, PYTHONPATH=ROOT, DETECTRON2_DATASETS=str(tmp_path / "d2"))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    with open(os.path.join(out, "tiny", "results.pkl"), "rb") as f:
        res = pickle.load(f)
    for ds in ("road_anomaly", "fishyscapes_laf"):
        for k in ("auroc", "aupr", "fpr95"):
            v = float(res[ds][k])
            assert 0.0 <= v <= 1.0, (ds, k, v)


def test_standin_cfgnode_and_registry():
    """The stand-in pieces the eval path executes, without the reference: CfgNode semantics the reference relies on
    (config.py assigns into detectron2's default tree; evaluate_ood.py:112-116 passes opts), registry override."""
    cfg = standins.get_cfg()
    standins.add_deeplab_config(cfg)
    cfg.MODEL.MASK_FORMER = standins.CfgNode()
    cfg.MODEL.MASK_FORMER.DEC_LAYERS = 10
    cfg.merge_from_list(["OUTPUT_DIR", "output/", "MODEL.MASK_FORMER.DEC_LAYERS", "2"])
    assert cfg.OUTPUT_DIR == "output/" and cfg.MODEL.MASK_FORMER.DEC_LAYERS == 2
    c2 = cfg.clone()
    cfg.freeze()
    with pytest.raises(AttributeError):
        cfg.MODEL.DEVICE = "cpu"
    c2.MODEL.DEVICE = "cpu"
    assert yaml.safe_load(c2.dump())["MODEL"]["DEVICE"] == "cpu"
    reg = standins.Registry("X")

    @reg.register()
    class A:          # noqa: N801
        pass
    assert reg.get("A") is A
    with pytest.raises(KeyError):
        reg.get("B")
    e = standins.EasyDict({"a": {"b": 1}, "eval-only": True})
    assert e.a.b == 1 and e["eval-only"] is True
    out = standins._ACompose([standins._AResize(8, 6), standins._AToTensorV2()])(
        image=np.zeros((4, 4, 3), np.uint8), mask=np.ones((4, 4), np.uint8))
    assert out["image"].shape == (3, 8, 6) and out["image"].dtype == torch.uint8 and out["mask"].shape == (8, 6)
    assert abs(standins.fpr_at_95_tpr(np.array([0.1, 0.2, 0.8, 0.9]), np.array([0, 0, 1, 1]))) < 1e-12


def test_plug_in_serves_maskformer_meta_arch():
    served = compat.plug_in()
    import detectron2.modeling as dm          # stand-in or real
    import rba_b200
    if "detectron2" in served:
        assert standins.PLUGIN_META_ARCH["MaskFormer"] is rba_b200.MaskFormer
        import detectron2.evaluation as de    # import-only placeholder package
        class E(de.DatasetEvaluator):         # noqa: N801  (subclassable)
            pass
        assert E is not None
    assert hasattr(dm, "build_model")
    import MultiScaleDeformableAttention as MSDA
    assert hasattr(MSDA, "ms_deform_attn_forward")


@pytest.mark.gpu
@needs_ref
def test_evaluate_ood_runs_unchanged_on_gpu(tmp_path):
    """The script itself (evaluate_ood.py:195-288), unchanged, with the REAL engine serving the model on the B200:
    `cd <reference> && python -m rba_b200.compat.run evaluate_ood.py ...`.  On the GPU box the reference
    files come from baseline/_ref.  Metrics are checked against this repo's device-resident evaluator on the same data."""
    data, models, out = str(tmp_path / "data"), str(tmp_path / "models"), str(tmp_path / "results")
    make_dataset_tree(data, h=128, w=192, n_laf=3)
    mc, sd = make_models_folder(models, "cuda")
    cmd = [sys.executable, "-m", "rba_b200.compat.run", "evaluate_ood.py", "--datasets_folder", data, "--models_folder", models,
           "--out_path", out, "--num_workers", "0", "--device", "cuda"]
    env = dict(os.environ, PYTHONPATH=ROOT, DETECTRON2_DATASETS=str(tmp_path / "d2"))
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REF, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    with open(os.path.join(out, "tiny", "results.pkl"), "rb") as f:
        res = pickle.load(f)
    for ds in ("road_anomaly", "fishyscapes_laf"):
        for k in ("auroc", "aupr", "fpr95"):
            assert 0.0 <= float(res[ds][k]) <= 1.0, (ds, k)
    # the same numbers from the fused score + device-resident metrics (rba_b200.OODEvaluator)
    import cv2
    import rba_b200
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(sd)
    model.to("cuda").eval()
    fs = os.path.join(data, "Fishyscapes")
    items = []
    for lbl in sorted(os.listdir(os.path.join(fs, "fishyscapes_lostandfound"))):
        img = cv2.imread(os.path.join(fs, "laf_images", lbl[5:-10] + "leftImg8bit.png"))[:, :, ::-1].copy()
        lab = cv2.imread(os.path.join(fs, "fishyscapes_lostandfound", lbl))[:, :, 0]
        items.append((torch.from_numpy(img).permute(2, 0, 1).contiguous(), torch.from_numpy(lab.copy()).long()))
    m = rba_b200.OODEvaluator(model).evaluate_dataset(items, batch=2, workers=1)
    for k in ("auroc", "aupr"):
        assert abs(m[k] - float(res["fishyscapes_laf"][k])) < 2e-3, (k, m[k], float(res["fishyscapes_laf"][k]))
