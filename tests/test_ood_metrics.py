"""OoD metrics.  CPU: the sklearn restatement (oracle/ood_metrics_oracle.py) is pinned against the
golden outputs of the reference's own OODEvaluator (support.py:247-303).  GPU: the histogram kernels
(csrc/ood_metrics.cu) through the C ABI against that oracle — exact on quantised scores, 1e-4 on raw scores."""
import numpy as np
import pytest
import torch

import ood_metrics_oracle as MO
from conftest import load_golden

METRIC_TOL = 1e-4      # raw scores: quantisation to 2^-15 relative resolution merges a handful of thresholds
KEYS = ("auroc", "aupr", "fpr95")


def test_metrics_oracle_matches_reference_golden():
    fix = load_golden("ood_metrics.pt")
    for nm, f in fix.items():
        got = MO.evaluate_ood(f["score"].numpy(), f["gt"].numpy())
        for k in KEYS:
            assert abs(got[k] - f["metrics"][k]) < 1e-12, (nm, k, got[k], f["metrics"][k])


def test_quantize_like_kernel_is_monotone_and_tight():
    rng = np.random.default_rng(0)
    s = np.concatenate([rng.standard_normal(10000).astype(np.float32) * 10, np.float32([0.0, -0.0, 1e-30, -1e-30, 19.0, -19.0])])
    q = MO.quantize_like_kernel(s)
    o = np.argsort(s, kind="stable")
    assert (np.diff(q[o]) >= 0).all()                                   # order preserved (ties allowed)
    nz = np.abs(s) > 1e-20
    assert (np.abs(q[nz] - s[nz]) <= np.abs(s[nz]) * 2.0 ** -15).all()  # 15 mantissa bits kept


@pytest.mark.gpu
def test_gpu_metrics_match_reference_golden(dev):
    import rba_b200
    fix = load_golden("ood_metrics.pt")
    for nm, f in fix.items():
        got = rba_b200.evaluate_ood(f["score"], f["gt"], device=dev)
        for k in KEYS:
            assert abs(got[k] - f["metrics"][k]) < METRIC_TOL, (nm, k, got[k], f["metrics"][k])
        # exact against sklearn on the quantised scores
        exact = MO.evaluate_ood(MO.quantize_like_kernel(f["score"].numpy()), f["gt"].numpy())
        for k in KEYS:
            assert abs(got[k] - exact[k]) < 1e-9, (nm, k, got[k], exact[k])


@pytest.mark.gpu
def test_gpu_metrics_streaming_and_edges(dev):
    import rba_b200
    g = torch.Generator().manual_seed(3)
    n = 1_000_003                                                        # not a multiple of the warp size
    gt = torch.randint(0, 3, (n,), generator=g)
    gt[gt == 2] = 255                                                    # ignored
    score = torch.randn(n, generator=g) * 2 - 17 + 3.0 * (gt == 1)
    score[::7] = -19.0                                                   # a heavy tie (saturated RbA)
    exact = MO.evaluate_ood(MO.quantize_like_kernel(score.numpy()), gt.numpy())
    raw = MO.evaluate_ood(score.numpy(), gt.numpy())
    m = rba_b200.StreamingOODMetrics(dev)
    m.update(score.to(dev), gt.to(dev))
    one = m.compute()
    assert one["n_ood"] == int((gt == 1).sum()) and one["n_ind"] == int((gt == 0).sum())
    for k in KEYS:
        assert abs(one[k] - exact[k]) < 1e-9 and abs(one[k] - raw[k]) < METRIC_TOL, (k, one[k], exact[k], raw[k])
    # streaming in ragged chunks with uint8 labels == one shot; result is deterministic bit for bit
    m2 = rba_b200.StreamingOODMetrics(dev)
    for lo, hi in [(0, 17), (17, 400_000), (400_000, 400_000), (400_000, n)]:
        m2.update(score[lo:hi].to(dev), gt[lo:hi].to(torch.uint8).to(dev))
    two = m2.compute()
    assert two == one and m.compute() == one
    # no positives / no negatives -> NaN, like an undefined ROC
    m.reset()
    m.update(score[:1000].to(dev), torch.zeros(1000, dtype=torch.int64, device=dev))
    r = m.compute()
    assert np.isnan(r["auroc"]) and np.isnan(r["aupr"]) and r["n_ood"] == 0
    # NaN scores are skipped, mismatched sizes raise
    m.reset()
    s3 = torch.tensor([0.1, float("nan"), 0.9, 0.2], device=dev)
    m.update(s3, torch.tensor([0, 1, 1, 0], device=dev))
    r = m.compute()
    assert r["n_ood"] == 1 and r["auroc"] == 1.0
    with pytest.raises(rba_b200.RbaError):
        m.update(s3, torch.zeros(3, dtype=torch.int64, device=dev))


@pytest.mark.gpu
def test_gpu_ood_evaluator_loop(dev):
    """rba_b200.OODEvaluator: the reference's compute_anomaly_scores + evaluate_ood flow (support.py:353-399, 270-303)
    with scores kept on the device, against the oracle metrics of the model's own score maps."""
    import rba_b200
    from golden_cases import CASES, case_model_config
    from rba_b200 import weights
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"]))
    model.to(dev).eval()
    g = torch.Generator().manual_seed(9)
    xs = [torch.randint(0, 256, (1, 3, 64, 96), dtype=torch.uint8, generator=g) for _ in range(3)]
    ys = [torch.randint(0, 2, (1, 64, 96), generator=g) for _ in range(3)]
    res = rba_b200.OODEvaluator(model).evaluate(list(zip(xs, ys)), upper_limit=2)     # third batch is cut off
    maps = torch.cat([model.rba([{"image": x[0].to(dev)}]) for x in xs[:2]]).cpu().numpy()
    exact = MO.evaluate_ood(MO.quantize_like_kernel(maps), torch.cat(ys[:2]).numpy())
    for k in KEYS:
        assert abs(res[k] - exact[k]) < 1e-9


@pytest.mark.gpu
def test_gpu_ood_evaluator_pipelined_dataset(dev):
    """OODEvaluator.evaluate_dataset (PinnedBatcher threads -> pinned batches -> ScoreStream(d2h=False) -> device
    histogram) gives the same metrics as the synchronous per-image loop, including a padded last batch and HWC input."""
    import rba_b200
    from golden_cases import CASES, case_model_config
    from rba_b200 import weights
    case = CASES["tiny_1dl"]
    mc = case_model_config(case)
    model = rba_b200.MaskFormer(mc)
    model.load_state_dict(weights.init_state_dict(mc, seed=case["seed"], perturb=case["perturb"]))
    model.to(dev).eval()

    class DS:
        def __len__(self):
            return 7

        def __getitem__(self, i):
            g = torch.Generator().manual_seed(100 + i)
            img = torch.randint(0, 256, (64, 96, 3), dtype=torch.uint8, generator=g).numpy()     # HWC like cv2
            lab = torch.randint(0, 3, (64, 96), generator=g)
            lab[lab == 2] = 255                                                                  # ignored region
            return img, lab.numpy()

    ds = DS()
    ev = rba_b200.OODEvaluator(model)
    loop = ev.evaluate([(torch.as_tensor(ds[i][0]).permute(2, 0, 1)[None], torch.as_tensor(ds[i][1])[None]) for i in range(7)],
                       upper_limit=7)
    for use_graph in (True, False):
        piped = ev.evaluate_dataset(ds, batch=3, workers=4, use_graph=use_graph)
        for k in KEYS:
            assert abs(piped[k] - loop[k]) < 1e-12, (k, piped[k], loop[k])
    part = ev.evaluate_dataset(ds, batch=2, workers=2, upper_limit=4)
    loop4 = ev.evaluate([(torch.as_tensor(ds[i][0]).permute(2, 0, 1)[None], torch.as_tensor(ds[i][1])[None]) for i in range(4)],
                        upper_limit=4)
    for k in KEYS:
        assert abs(part[k] - loop4[k]) < 1e-12
