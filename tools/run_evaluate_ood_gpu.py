"""BASELINE config 5 / VERDICT r1 item 8: the reference's evaluate_ood.py, UNCHANGED, on the B200 through
rba_b200.compat (stand-ins for detectron2 & co., META_ARCH "MaskFormer" -> rba_b200.MaskFormer), on a synthetic
FS-LaF-shaped tree (1024 x 2048 PNGs) -- then the same images through this repo's own pipelined evaluator
(OODEvaluator.evaluate_dataset) and, with --gpus N, sharded over N processes with one gather of the score maps.

    python tools/run_evaluate_ood_gpu.py [--arch swin_b_1dl|swin_b_full] [--images 16] [--out gpurun_out/evaluate_ood.json]

evaluate_ood.py comes from baseline/_ref (tools/make_baseline_ref.py; the GPU box has no /root/reference)."""
import argparse
import json
import os
import pickle
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arch", default="swin_b_1dl")
    ap.add_argument("--images", type=int, default=16)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "evaluate_ood.json"))
    args = ap.parse_args()
    import torch
    import test_compat_evaluate_ood as T
    ref = T.REF
    assert os.path.isfile(os.path.join(ref, "evaluate_ood.py")), "no reference checkout / baseline/_ref"
    res = {"arch": args.arch, "images": args.images, "hw": [args.height, args.width], "reference_root": ref}
    with tempfile.TemporaryDirectory() as tmp:
        data, models, out = os.path.join(tmp, "data"), os.path.join(tmp, "models"), os.path.join(tmp, "results")
        T.make_dataset_tree(data, h=args.height, w=args.width, n_laf=args.images, ra_hw=(720, 1280))
        T.make_models_folder(models, "cuda", arch=args.arch)
        cmd = [sys.executable, "-m", "rba_b200.compat.run", "evaluate_ood.py", "--datasets_folder", data, "--models_folder",
               models, "--out_path", out, "--num_workers", "4", "--device", "cuda"]
        env = dict(os.environ, PYTHONPATH=ROOT, DETECTRON2_DATASETS=os.path.join(tmp, "d2"))
        t0 = time.time()
        r = subprocess.run(cmd, capture_output=True, text=True, cwd=ref, env=env, timeout=3000)
        wall = time.time() - t0
        if r.returncode != 0:
            sys.stderr.write(r.stdout[-3000:] + "\n" + r.stderr[-6000:])
            raise SystemExit("evaluate_ood.py failed")
        with open(os.path.join(out, args.arch, "results.pkl"), "rb") as f:
            rec = pickle.load(f)
        res["evaluate_ood_py_unchanged"] = {
            "wall_s": wall, "images_scored": args.images + 2,
            "images_per_s_whole_script": (args.images + 2) / wall,
            "results": {ds: {k: float(v) for k, v in rec[ds].items()} for ds in rec},
            "note": "whole-script wall clock: imports, model build + weight load, PNG decode (DataLoader, batch 1), forward, "
                    "per-image .cpu().numpy(), sklearn metrics over all pixels"}
        # ---- the same FS-LaF images through this repo's pipelined evaluator (device-resident metrics) ----
        import rba_b200
        from rba_b200 import compat
        compat.plug_in()
        sys.path.insert(0, ref)
        from rba_b200.compat.run import prefer_local_namespace_packages
        prefer_local_namespace_packages(ref)
        from datasets.fishyscapes import FishyscapesLAF       # the reference's own dataset class
        import albumentations as A
        from albumentations.pytorch import ToTensorV2
        ds = FishyscapesLAF(hparams=__import__("easydict").EasyDict(dataset_root=os.path.join(data, "Fishyscapes")),
                            transforms=A.Compose([ToTensorV2()]))
        import yaml
        from rba_b200.config import model_config_from_cfg
        mc = model_config_from_cfg(yaml.safe_load(open(os.path.join(models, args.arch, "config.yaml"))))
        model = rba_b200.MaskFormer(mc)
        model.load_state_dict(torch.load(os.path.join(models, args.arch, "model_final.pth"))["model"])
        model.to("cuda").eval()
        ev = rba_b200.OODEvaluator(model)
        ev.evaluate_dataset(ds, batch=8, workers=8)          # warm-up (capture, page cache)
        torch.cuda.synchronize()
        t0 = time.time()
        m = ev.evaluate_dataset(ds, batch=8, workers=8)
        torch.cuda.synchronize()
        dt = time.time() - t0
        res["pipelined_evaluator"] = {"wall_s": dt, "images_per_s": len(ds) / dt, "results": m,
                                      "reference_results": res["evaluate_ood_py_unchanged"]["results"].get("fishyscapes_laf"),
                                      "note": "PinnedBatcher (threaded PNG decode) + ScoreStream + device-resident AUROC/AUPR/FPR95"}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
