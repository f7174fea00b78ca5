"""tests/golden/panoptic.pt: the reference's own MaskFormer.panoptic_inference (mask2former/maskformer_model.py:394-481, imported
UNMODIFIED, called as an unbound method on a stand-in `self` carrying the six attributes it reads) on seeded synthetic head
outputs.  TEST INFRASTRUCTURE; run manually in the build container:  python oracle/make_golden_panoptic.py"""
import os
import sys
from types import SimpleNamespace

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_loader  # noqa: E402

CASES = {
    # name: Q, K, H, W, seed, mask logit shift (controls how many pixels each query claims), open_panoptic, threshold, pixel_min
    "closed": dict(Q=100, K=19, H=48, W=80, seed=1, shift=0.0, open=False, thr=-0.1, pmin=300, active=10),
    "open": dict(Q=100, K=19, H=64, W=96, seed=2, shift=0.0, open=True, thr=-0.3, pmin=20, active=12),
    "open_ret": dict(Q=40, K=19, H=40, W=56, seed=3, shift=0.0, open=True, thr=-0.3, pmin=5, active=8),
    "nothing_kept": dict(Q=6, K=19, H=16, W=16, seed=4, shift=0.0, open=True, thr=-0.1, pmin=3, void=True, active=1),
}
THINGS = [11, 12, 13, 14, 15, 16, 17, 18]         # cityscapes thing classes (contiguous ids)


def make_inputs(c):
    """Head outputs with structure: `active` queries own disjoint blobs of a random low-resolution partition (mask logit +6 inside,
    -6 outside, plus noise), carry confident class logits (classes repeat, so stuff regions merge; some are things), one of
    them overlaps a neighbour (overlap filter), the rest predict void with empty masks; region 0 of the partition is claimed by
    nobody (high RbA score there: the open-panoptic branch finds it)."""
    g = torch.Generator().manual_seed(c["seed"])
    Q, K, H, W = c["Q"], c["K"], c["H"], c["W"]
    active = c.get("active", 10)
    part = torch.rand(active + 1, H // 8, W // 8, generator=g)
    part = torch.nn.functional.interpolate(part[None], size=(H, W), mode="bilinear", align_corners=False)[0].argmax(0)   # (H,W) in 0..active
    masks = torch.full((Q, H, W), -6.0) + torch.randn(Q, H, W, generator=g) * 0.3
    cls = torch.randn(Q, K + 1, generator=g)
    cls[:, -1] += 12.0                                        # default: void
    classes = [2, 2, 8, 11, 13, 0, 10, 13, 5, 2, 17, 1][:active]
    for k in range(active):
        q = 3 * k + 1
        masks[q][part == k + 1] += 12.0
        cls[q] = torch.randn(K + 1, generator=g)
        cls[q, classes[k]] += 12.0
    if active >= 2 and not c.get("void"):                     # a confident query that mostly duplicates query 1's region but scores lower
        masks[0][part == 1] += 12.0
        masks[0][part == 2] += 12.0
        cls[0] = torch.randn(K + 1, generator=g)
        cls[0, 7] += 6.0
    if c.get("void"):
        cls[:, -1] += 30.0
    return cls, masks + c["shift"]


def main():
    ref_loader._install()
    from mask2former.maskformer_model import MaskFormer
    fix = {}
    for name, c in CASES.items():
        cls, masks = make_inputs(c)
        fake = SimpleNamespace(sem_seg_head=SimpleNamespace(num_classes=c["K"]), object_mask_threshold=0.8, overlap_threshold=0.8,
                               metadata=SimpleNamespace(thing_dataset_id_to_contiguous_id={i: t for i, t in enumerate(THINGS)}),
                               device=torch.device("cpu"))
        out = MaskFormer.panoptic_inference(fake, cls, masks, c["open"], c["thr"], c["pmin"], name == "open_ret")
        fix[name] = {"case": c, "panoptic_seg": out[0].clone(), "segments_info": out[1], "in_checksum": float(cls.double().sum() + masks.double().sum())}
        if name == "open_ret":
            fix[name]["ood_mask"] = out[2].clone()
        print(name, "segments", len(out[1]), "ids", sorted(set(out[0].flatten().tolist()))[:12])
    torch.save({"cases": fix, "things": THINGS}, os.path.join(os.path.dirname(HERE), "tests", "golden", "panoptic.pt"))


if __name__ == "__main__":
    main()
