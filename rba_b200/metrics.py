"""GPU-resident OoD evaluation: the reference's `OODEvaluator` (support.py:219-399) restated so that
score maps never leave the device and the metrics never touch sklearn.

    ev = rba_b200.OODEvaluator(model)                    # model: rba_b200.MaskFormer on CUDA
    res = ev.evaluate(loader, upper_limit=1300)          # {'auroc': .., 'aupr': .., 'fpr95': ..}

`StreamingOODMetrics` is the accumulator underneath: `update(score, gt)` adds a batch of (score, label) pixels to a
two-class 2^24-bin histogram of order-preserving float keys (csrc/ood_metrics.cu), `compute()` sweeps it.  The
result equals sklearn's roc_curve/auc/average_precision_score on scores quantised to 2^-15 relative resolution
(exactly), and the reference's numbers within ~1e-5 on real score maps.

One documented difference: `fpr95` is the FPR at the FIRST threshold whose TPR exceeds 0.95.  The reference's
`calculate_auroc` (support.py:247-268) walks sklearn's `roc_curve` output with its default `drop_intermediate=True`, which
removes collinear ROC points, so the first RETAINED point with TPR > 0.95 can be a later one with a slightly higher FPR.  AUROC
and AUPR are unaffected; on the committed goldens all three metrics agree with the reference's within the tolerance of tests/test_ood_metrics.py."""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import RbaError


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class StreamingOODMetrics:
    def __init__(self, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RbaError("StreamingOODMetrics runs on CUDA only (no CPU fallback)")
        L = _lib.lib()
        self._hist = torch.zeros(int(L.rba_ood_hist_bytes()) // 8, dtype=torch.int64, device=self.device)
        self._ws = torch.empty(int(L.rba_ood_workspace_bytes()) // 8, dtype=torch.int64, device=self.device)
        self._out = torch.empty(5, dtype=torch.float64, device=self.device)

    def reset(self):
        self._hist.zero_()

    def update(self, score, gt):
        """score: float32 CUDA tensor of any shape; gt: same number of elements, uint8 or int64 (1 = OoD, 0 =
        in-distribution, anything else ignored — support.py:275-279)."""
        if not score.is_cuda or not gt.is_cuda:
            raise RbaError("StreamingOODMetrics.update: CUDA tensors expected")
        if score.numel() != gt.numel():
            raise RbaError(f"score has {score.numel()} elements, gt {gt.numel()}")
        score = score.contiguous()
        if score.dtype != torch.float32:
            score = score.float()
        if gt.dtype not in (torch.uint8, torch.int64):
            gt = gt.to(torch.int64)
        gt = gt.contiguous()
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().rba_ood_hist_update(_p(score), _p(gt), gt.element_size(), score.numel(), _p(self._hist), st))

    def compute(self):
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(_lib.lib().rba_ood_hist_finalize(_p(self._hist), _p(self._ws), _p(self._out), st))
        o = self._out.cpu().tolist()               # the only D2H of the whole evaluation: 40 bytes
        return {"auroc": o[0], "aupr": o[1], "fpr95": o[2], "n_ood": int(o[3]), "n_ind": int(o[4])}


def evaluate_ood(anomaly_score, ood_gts, device="cuda"):
    """Same arguments and result keys as OODEvaluator.evaluate_ood (support.py:270-303): anomaly_score and ood_gts are
    arrays / tensors of equal size."""
    m = StreamingOODMetrics(device)
    s = torch.as_tensor(np.asarray(anomaly_score) if not torch.is_tensor(anomaly_score) else anomaly_score)
    g = torch.as_tensor(np.asarray(ood_gts) if not torch.is_tensor(ood_gts) else ood_gts)
    m.update(s.to(m.device, torch.float32), g.to(m.device))
    r = m.compute()
    return {k: r[k] for k in ("auroc", "aupr", "fpr95")}


class OODEvaluator:
    """Mirror of the reference's OODEvaluator for the rba_b200 model: same loop as compute_anomaly_scores
    (support.py:353-399: x, y batches from a DataLoader, `upper_limit` images) but the score is the fused kernel's
    output and accumulation happens on the device."""

    def __init__(self, model, score_func="rba"):
        self.model = model
        self.score_func = score_func

    @torch.no_grad()
    def evaluate(self, loader, upper_limit=450, metrics=None):
        dev = self.model.device
        metrics = metrics or StreamingOODMetrics(dev)
        for jj, (x, y) in enumerate(loader):
            if jj >= upper_limit:
                break
            score = self.model.score([{"image": im} for im in x], self.score_func)       # (B,H,W) on the device
            metrics.update(score, y.to(dev, non_blocking=True))
        r = metrics.compute()
        return {k: r[k] for k in ("auroc", "aupr", "fpr95")}

    @torch.no_grad()
    def evaluate_dataset(self, dataset, batch=8, workers=8, upper_limit=None, metrics=None, use_graph=True, group=None):
        """The same evaluation over an indexable dataset of (image, label) pairs (the reference's dataset classes), fully
        pipelined: threaded decode into pinned batches (PinnedBatcher), H2D of batch i+1 overlapping the forward of batch i
        (ScoreStream with d2h=False), scores and labels accumulated on the device.  One 40-byte D2H at the end.

        Multi-GPU (one process per GPU, torch.distributed initialised): image i goes to rank i % world, every rank fills its
        own histogram and ONE all-reduce of the 2 x 2^24 counters at the end gives every rank the metrics over all pixels --
        no score map ever crosses the link (the reference gathers every map to the host, support.py:353-399)."""
        import torch.distributed as dist
        from .pipeline import PinnedBatcher, ScoreStream
        dev = self.model.device
        metrics = metrics or StreamingOODMetrics(dev)
        n = len(dataset) if upper_limit is None else min(len(dataset), int(upper_limit))
        world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank(group) if world > 1 else 0
        mine = range(rank, n, world)
        eng = self.model.engine()
        eng.set_score(self.score_func, include_void=False)
        stream, ybuf, ev_lab = None, None, None
        comp = torch.cuda.current_stream(dev)
        for it, (images, labels, n_valid) in enumerate(PinnedBatcher(dataset, batch, indices=mine, workers=workers)):
            if stream is None:
                stream = ScoreStream(eng, images.shape[0], images.shape[2], images.shape[3], use_graph=use_graph, d2h=False)
                ybuf = [torch.empty(labels.shape, dtype=torch.uint8, device=dev) for _ in range(2)]
                ev_lab = [torch.cuda.Event(), torch.cuda.Event()]
                ev_used = [torch.cuda.Event(), torch.cuda.Event()]
            k = it & 1
            with torch.cuda.stream(stream.s_in):              # labels ride the copy-in stream next to the images
                if it >= 2:
                    stream.s_in.wait_event(ev_used[k])
                ybuf[k].copy_(labels, non_blocking=True)
                ev_lab[k].record(stream.s_in)
            score = stream.step(images)                       # device tensor of this batch (stream-ordered)
            comp.wait_event(ev_lab[k])
            metrics.update(score, ybuf[k])
            ev_used[k].record(comp)
            # host: wait only for the two H2D copies (about a millisecond) so the pinned buffers can go back to the
            # ring; the forward keeps running while the next batch is fetched
            stream.ev_in_ready[k].synchronize()
            ev_lab[k].synchronize()
        if world > 1:
            dist.all_reduce(metrics._hist, group=group)           # int64 counters: exact, order-independent
        r = metrics.compute()
        return {k: r[k] for k in ("auroc", "aupr", "fpr95")}
