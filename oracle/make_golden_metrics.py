"""Generates tests/golden/ood_metrics.pt with the UNMODIFIED reference OODEvaluator (support.py) run in the build
container.  TEST INFRASTRUCTURE; run manually:  python oracle/make_golden_metrics.py
support.py imports albumentations / easydict / ood_metrics / detectron2-free code only; the missing third parties are
served by rba_b200.compat stand-ins (import-only here: evaluate_ood is pure numpy + sklearn)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
REF = os.environ.get("RBA_REFERENCE_ROOT", "/root/reference")

from rba_b200 import compat  # noqa: E402
from rba_b200.compat.run import prefer_local_namespace_packages  # noqa: E402

compat.plug_in()
sys.path.insert(0, REF)
prefer_local_namespace_packages(REF)
from support import OODEvaluator  # noqa: E402

ev = OODEvaluator(None, None, None)
rng = np.random.default_rng(0)
cases = {}
# (a) RbA-like saturated scores: in-distribution near -K.., OoD higher, heavy ties; labels with ignore regions
n_img, H, W = 3, 48, 64
gt = rng.choice([0, 1, 255], size=(n_img, 1, H, W), p=[0.85, 0.05, 0.10]).astype(np.int64)
score = (-18.5 + 2.5 * rng.random((n_img, H, W)) + 6.0 * (gt[:, 0] == 1) * rng.random((n_img, H, W))).astype(np.float32)
score = np.round(score, 2)                      # ties
cases["rba_like"] = (score, gt)
# (b) barely separable gaussian scores, all pixels labelled
gt2 = (rng.random((2, 1, 40, 40)) < 0.3).astype(np.int64)
score2 = (rng.standard_normal((2, 40, 40)) + 0.4 * gt2[:, 0]).astype(np.float32)
cases["gaussian"] = (score2, gt2)
# (c) perfectly separable and fully inverted
gt3 = np.zeros((1, 1, 8, 8), np.int64); gt3[0, 0, :2] = 1
score3 = gt3[:, 0].astype(np.float32) * 2 - 1
cases["separable"] = (score3, gt3)
cases["inverted"] = (-score3, gt3)
fix = {}
for nm, (s, g) in cases.items():
    res = ev.evaluate_ood(anomaly_score=s, ood_gts=g, verbose=False)
    fix[nm] = {"score": torch.from_numpy(s), "gt": torch.from_numpy(g), "metrics": {k: float(v) for k, v in res.items()}}
    print(nm, fix[nm]["metrics"])
torch.save(fix, os.path.join(ROOT, "tests", "golden", "ood_metrics.pt"))
