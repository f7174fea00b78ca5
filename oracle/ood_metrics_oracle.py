"""CPU restatement of the reference's OoD metric arithmetic (TEST INFRASTRUCTURE; never imported by the product).

Follows support.py:247-303 line by line: sklearn.metrics.roc_curve + auc for AUROC, the first ROC point with
tpr > 0.95 for FPR@95 (calculate_auroc, :247-257), average_precision_score for AUPR (:259-268), over the pixels whose
label is 0 (in-distribution) or 1 (OoD) (evaluate_ood, :270-303).  Pinned against the reference's own OODEvaluator by
tests/golden/ood_metrics.pt (oracle/make_golden_metrics.py)."""
import numpy as np
from sklearn.metrics import auc, average_precision_score, roc_curve


def calculate_auroc(conf, gt):
    """support.py:247-257"""
    fpr, tpr, threshold = roc_curve(gt, conf)
    roc_auc = auc(fpr, tpr)
    fpr_best = 0
    for i, j, k in zip(tpr, fpr, threshold):
        if i > 0.95:
            fpr_best = j
            break
    return roc_auc, fpr_best


def evaluate_ood(anomaly_score, ood_gts):
    """support.py:270-303: anomaly_score, ood_gts arrays of equal shape -> {'auroc','aupr','fpr95'}."""
    ood_gts = np.asarray(ood_gts).squeeze()
    anomaly_score = np.asarray(anomaly_score).squeeze()
    ood_out = anomaly_score[ood_gts == 1]
    ind_out = anomaly_score[ood_gts == 0]
    val_out = np.concatenate((ind_out, ood_out))
    val_label = np.concatenate((np.zeros(len(ind_out)), np.ones(len(ood_out))))
    prc_auc = average_precision_score(val_label, val_out)          # support.py:263
    roc_auc, fpr = calculate_auroc(val_out, val_label)             # support.py:264
    return {"auroc": float(roc_auc), "aupr": float(prc_auc), "fpr95": float(fpr)}


def quantize_like_kernel(score, bits=24):
    """The value every score of a histogram bin stands for (lower edge of the bin of the order-preserving key):
    sklearn on quantised scores is what rba_ood_hist_finalize computes exactly."""
    s = np.ascontiguousarray(score, dtype=np.float32) + np.float32(0.0)     # -0.0 -> +0.0 like the kernel
    u = s.view(np.uint32)
    key = np.where(u & 0x80000000, ~u, u | 0x80000000).astype(np.uint32)
    key = (key >> (32 - bits)) << (32 - bits)
    back = np.where(key & 0x80000000, key & 0x7FFFFFFF, ~key).astype(np.uint32)
    return back.view(np.float32)
