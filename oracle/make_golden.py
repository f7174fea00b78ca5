"""Generates tests/golden/*.pt by running the UNMODIFIED reference (imported from /root/reference via
oracle/ref_loader.py) in the build container.  TEST INFRASTRUCTURE; run manually:

    python oracle/make_golden.py

Fixtures are small: weights are not stored, they are re-generated from (config preset, seed, perturb) by
rba_b200.weights.init_state_dict and guarded by a checksum; stored are the inputs' seeds and the reference's
outputs (pred_logits, pred_masks, rba, sem_seg subsampled by 4).
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
warnings.filterwarnings("ignore")

import ref_loader  # noqa: E402
from golden_cases import CASES, case_model_config, case_images, state_checksum  # noqa: E402
from rba_b200 import weights  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_overrides(case):
    o = {}
    if case["preset"].startswith("tiny"):
        o.update({"MODEL.SWIN.EMBED_DIM": 32, "MODEL.SWIN.DEPTHS": [2, 2, 2, 2], "MODEL.SWIN.NUM_HEADS": [1, 2, 4, 8],
                  "MODEL.SEM_SEG_HEAD.TRANSFORMER_ENC_LAYERS": 2})
    if case.get("levels", 1) == 3:
        o["MODEL.SEM_SEG_HEAD.DEFORMABLE_TRANSFORMER_ENCODER_IN_FEATURES"] = ["res3", "res4", "res5"]
    o["MODEL.MASK_FORMER.DEC_LAYERS"] = case.get("dec_layers", 1) + 1
    if case.get("ood_prediction"):
        o["MODEL.MASK_FORMER.DENSE_HYBRID_LOSS"] = True
    return o


@torch.no_grad()
def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    only = set(sys.argv[1:])            # e.g. `python oracle/make_golden.py tiny_ood`: just that model case
    for name, case in CASES.items():
        if only and name not in only:
            continue
        base = "swin_l_1dl" if case["preset"] == "swin_l_1dl" else "swin_b_1dl"
        cfg = ref_loader.load_cfg(base, reference_overrides(case))
        model = ref_loader.build_reference_model(cfg, seed=0)
        mc = case_model_config(case)
        sd = weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"])
        ref_sd = model.state_dict()
        assert list(ref_sd.keys()) == list(sd.keys()), "state_dict keys differ from the reference"
        sd_meta = type(ref_sd)(sd)
        sd_meta._metadata = ref_sd._metadata
        model.load_state_dict(sd_meta)
        images = case_images(case)
        caps = {}
        model.sem_seg_head.register_forward_hook(lambda m, i, o: caps.__setitem__("head", o))
        if case.get("ood_prediction"):
            out, ood_pred = model([{"image": im} for im in images], return_ood_pred=True)
        else:
            out, ood_pred = model([{"image": im} for im in images]), None
        sem = torch.stack([o["sem_seg"] for o in out])
        rba = -sem.tanh().sum(1)                                     # evaluate_ood.py:148-150
        import rba_oracle as O
        margin = O.forward(sd, mc, images, want_taps=True)["taps"]["am_margin"]
        fix = {
            "case": case, "state_checksum": state_checksum(sd), "am_margin": margin,
            "pred_logits": caps["head"]["pred_logits"].clone(), "pred_masks": caps["head"]["pred_masks"].clone(),
            "rba": rba.clone(), "sem_seg_s4": sem[:, :, ::4, ::4].clone(),
            "torch_version": torch.__version__, "reference": "NazirNayal8/RbA @ /root/reference (unmodified modules under oracle/ref_shims)",
        }
        if ood_pred is not None:
            import torch.nn.functional as Fn
            # evaluate_ood.py:161-173 get_densehybrid_score, verbatim arithmetic
            p2 = Fn.softmax(ood_pred, dim=1)[:, 1]
            fix["ood_pred"] = ood_pred.clone()
            fix["densehybrid"] = (-torch.logsumexp(sem, dim=1)) + (p2 + 1e-9).log()
            fix["energy"] = -torch.logsumexp(sem, dim=1)
        torch.save(fix, os.path.join(OUT, f"model_{name}.pt"))
        print(name, "pred_masks", tuple(fix["pred_masks"].shape), "rba range", float(rba.min()), float(rba.max()), "am_margin", margin)

    if only:
        return
    # ---- MSDeformAttn: shapes / seed / value scaling of the reference's own ops/test.py:24-47 (CPU RNG) ----
    core = ref_loader.msda_core_pytorch()
    N, M, D, Lq, L, P = 1, 2, 2, 2, 2, 2
    shapes = torch.as_tensor([(6, 4), (3, 2)], dtype=torch.long)
    S = int((shapes[:, 0] * shapes[:, 1]).sum())
    torch.manual_seed(3)
    value = torch.rand(N, S, M, D) * 0.01
    loc = torch.rand(N, Lq, M, L, P, 2)
    aw = torch.rand(N, Lq, M, L, P) + 1e-5
    aw /= aw.sum(-1, keepdim=True).sum(-2, keepdim=True)
    fix = {"value": value, "shapes": shapes, "loc": loc, "aw": aw, "out": core(value, shapes, loc, aw),
           "out_double": core(value.double(), shapes, loc.double(), aw.double())}
    # a larger, out-of-range-heavy case (locations in [-0.3, 1.3]) at the model's real head layout M=8, D=32
    N, M, D, Lq, L, P = 2, 8, 32, 37, 3, 4
    shapes2 = torch.as_tensor([(5, 7), (10, 14), (3, 4)], dtype=torch.long)
    S2 = int((shapes2[:, 0] * shapes2[:, 1]).sum())
    value2 = torch.randn(N, S2, M, D)
    loc2 = torch.rand(N, Lq, M, L, P, 2) * 1.6 - 0.3
    aw2 = torch.softmax(torch.randn(N, Lq, M, L * P), -1).view(N, Lq, M, L, P)
    fix["big"] = {"value": value2, "shapes": shapes2, "loc": loc2, "aw": aw2, "out": core(value2, shapes2, loc2, aw2)}
    torch.save(fix, os.path.join(OUT, "msda.pt"))
    print("msda", tuple(fix["out"].shape), tuple(fix["big"]["out"].shape))

    # ---- fused score: the reference's own upsample + semantic_inference + get_RbA arithmetic ----
    from mask2former.maskformer_model import MaskFormer
    import torch.nn.functional as F
    torch.manual_seed(5)
    fixs = {}
    for nm, (B, Q, K, h, w, H, W) in {"a": (2, 100, 19, 6, 9, 24, 36), "crop": (1, 100, 19, 8, 8, 29, 30),
                                       "q7": (1, 7, 19, 5, 40, 20, 160)}.items():
        masks = torch.randn(B, Q, h, w) * 0.99 - 0.54          # stats of real mask logits
        logits = torch.randn(B, Q, K + 1)
        up = F.interpolate(masks, size=(4 * h, 4 * w), mode="bilinear", align_corners=False)   # maskformer_model.py:294-299
        sem = torch.stack([MaskFormer.semantic_inference(None, logits[b], up[b])[:, :H, :W] for b in range(B)])
        fixs[nm] = {"masks": masks, "logits": logits, "H": H, "W": W, "sem_seg": sem, "rba": -sem.tanh().sum(1)}
    torch.save(fixs, os.path.join(OUT, "score.pt"))
    print("score ok")


if __name__ == "__main__":
    main()
