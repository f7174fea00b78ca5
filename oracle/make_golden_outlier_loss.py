"""Generates tests/golden/outlier_loss.pt: loss value and input gradients of the reference's own
SetCriterion.outlier_loss (mask2former/modeling/criterion.py:435-553, imported UNMODIFIED from /root/reference and called
as an unbound function on a namespace carrying the five config attributes it reads), for the shipped configurations
(nls + tanh, energy; squared hinge) plus nls with sigmoid / no norm, with and without outlier pixels.  Inputs are
regenerated from seeds by the tests; stored are loss + gradients.  TEST INFRASTRUCTURE; run manually:
    python oracle/make_golden_outlier_loss.py
"""
import os
import sys
import types
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings("ignore")
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    "nls_tanh": dict(seed=1, B=2, Q=100, K=19, h=12, w=20, H=48, W=80, target="nls", norm="tanh", t_in=-1.0, t_out=-0.1, p_ood=0.2),
    "energy": dict(seed=2, B=1, Q=100, K=19, h=12, w=20, H=48, W=80, target="energy", norm="none", t_in=-1.0, t_out=-0.1, p_ood=0.2),
    "nls_sigmoid": dict(seed=3, B=1, Q=37, K=7, h=9, w=13, H=33, W=50, target="nls", norm="sigmoid", t_in=-4.0, t_out=-3.0, p_ood=0.3),
    "nls_none_no_ood": dict(seed=4, B=3, Q=10, K=19, h=16, w=16, H=64, W=64, target="nls", norm="none", t_in=-5.0, t_out=-0.5, p_ood=0.0),
    "nls_tanh_odd_ratio": dict(seed=5, B=1, Q=64, K=19, h=17, w=31, H=70, W=100, target="nls", norm="tanh", t_in=-6.0, t_out=-2.0, p_ood=0.1),
}


def make_inputs(c):
    """Same recipe in tests/test_kernels_gpu.py::_outlier_inputs."""
    g = torch.Generator().manual_seed(c["seed"])
    masks = torch.randn(c["B"], c["Q"], c["h"], c["w"], generator=g) * 0.99 - 0.54
    logits = torch.randn(c["B"], c["Q"], c["K"] + 1, generator=g)
    r = torch.rand(c["B"], c["H"], c["W"], generator=g)
    labels = torch.full((c["B"], c["H"], c["W"]), 255, dtype=torch.int64)      # ignored
    labels[r < 0.6] = 0
    labels[r > 1.0 - c["p_ood"]] = 1
    return masks, logits, labels


def main():
    ref_loader._install()
    from mask2former.modeling.criterion import SetCriterion
    fix = {}
    for name, c in CASES.items():
        masks, logits, labels = make_inputs(c)
        m, l = masks.double().requires_grad_(True), logits.double().requires_grad_(True)
        ns = types.SimpleNamespace(outlier_loss_target=c["target"], score_norm=c["norm"], outlier_loss_func="squared_hinge",
                                   inlier_upper_threshold=c["t_in"], outlier_lower_threshold=c["t_out"])
        targets = [{"outlier_masks": labels[b]} for b in range(c["B"])]
        loss = SetCriterion.outlier_loss(ns, {"pred_masks": m, "pred_logits": l}, targets, None, None)["outlier_loss"]
        loss.backward()
        fix[name] = {"case": c, "loss": float(loss), "d_masks": m.grad.float(), "d_logits": l.grad.float(),
                     "in_checksum": float(masks.double().sum() + logits.double().sum() + labels.double().sum())}
        print(name, float(loss), float(m.grad.abs().max()), float(l.grad.abs().max()))
    torch.save(fix, os.path.join(OUT, "outlier_loss.pt"))


if __name__ == "__main__":
    main()
