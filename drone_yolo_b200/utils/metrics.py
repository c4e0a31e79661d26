"""Detection metrics the validator caller needs (ultralytics/utils/metrics.py:67-95 `box_iou`, :447-452 `smooth`, :505-534
`compute_ap`, :537-623 `ap_per_class`, :626-760 `Metric`, :799-906 `DetMetrics`; engine/validator.py:224-264 `match_predictions`).

Host-side bookkeeping over the <= max_det rows per image that leave NMS: plain torch / numpy, not kernels.  The numerics
follow the reference step by step (same sort, same interpolation grids, same eps) so that the same detections give the
same numbers; plots and the confusion matrix are outside the hot path and are not built.
"""
from __future__ import annotations

import numpy as np
import torch


def box_iou(box1: torch.Tensor, box2: torch.Tensor, eps: float = 1e-7) -> torch.Tensor:
    """(N, 4) x (M, 4) xyxy -> (N, M) IoU (metrics.py:67-95): inter / (area1 + area2 - inter + eps) in the boxes' dtype."""
    a1, a2 = box1.float().unsqueeze(1).chunk(2, 2)
    b1, b2 = box2.float().unsqueeze(0).chunk(2, 2)
    inter = (torch.min(a2, b2) - torch.max(a1, b1)).clamp_(0).prod(2)
    return inter / ((a2 - a1).prod(2) + (b2 - b1).prod(2) - inter + eps)


def match_predictions(pred_classes: torch.Tensor, true_classes: torch.Tensor, iou: torch.Tensor, iouv: torch.Tensor) -> torch.Tensor:
    """(N,) predicted classes, (M,) label classes, (M, N) IoU -> (N, len(iouv)) bool "correct" matrix (validator.py:224-264).

    Per threshold: all (label, detection) pairs of equal class with IoU >= t, best IoU first; every detection keeps its best
    label, then every label keeps its first (best) detection."""
    correct = np.zeros((pred_classes.shape[0], iouv.shape[0]), dtype=bool)
    same = true_classes[:, None] == pred_classes
    iou = (iou * same).cpu().numpy()
    for i, t in enumerate(iouv.cpu().tolist()):
        pairs = np.array(np.nonzero(iou >= t)).T                     # rows (label, detection)
        if pairs.shape[0]:
            if pairs.shape[0] > 1:
                pairs = pairs[iou[pairs[:, 0], pairs[:, 1]].argsort()[::-1]]
                pairs = pairs[np.unique(pairs[:, 1], return_index=True)[1]]
                pairs = pairs[np.unique(pairs[:, 0], return_index=True)[1]]
            correct[pairs[:, 1].astype(int), i] = True
    return torch.tensor(correct, dtype=torch.bool, device=pred_classes.device)


def smooth(y: np.ndarray, f: float = 0.05) -> np.ndarray:
    """Box filter of fraction f (metrics.py:447-452)."""
    nf = round(len(y) * f * 2) // 2 + 1
    p = np.ones(nf // 2)
    return np.convolve(np.concatenate((p * y[0], y, p * y[-1]), 0), np.ones(nf) / nf, mode="valid")


def compute_ap(recall: np.ndarray, precision: np.ndarray):
    """101-point interpolated AP of one recall / precision curve (metrics.py:505-534)."""
    mrec = np.concatenate(([0.0], recall, [1.0]))
    mpre = np.concatenate(([1.0], precision, [0.0]))
    mpre = np.flip(np.maximum.accumulate(np.flip(mpre)))             # precision envelope
    x = np.linspace(0, 1, 101)
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return trapz(np.interp(x, mrec, mpre), x), mpre, mrec


def ap_per_class(tp: np.ndarray, conf: np.ndarray, pred_cls: np.ndarray, target_cls: np.ndarray, eps: float = 1e-16):
    """tp (D, 10) bool, conf (D,), pred_cls (D,), target_cls (L,) -> (p, r, f1, ap (nc_seen, 10), classes) at the max-F1
    confidence, as metrics.py:537-623 returns them (without the curves)."""
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, nt = np.unique(target_cls, return_counts=True)
    n = classes.shape[0]
    x = np.linspace(0, 1, 1000)
    ap, p_curve, r_curve = np.zeros((n, tp.shape[1])), np.zeros((n, 1000)), np.zeros((n, 1000))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        if sel.sum() == 0 or nt[ci] == 0:
            continue
        fpc = (1 - tp[sel]).cumsum(0)
        tpc = tp[sel].cumsum(0)
        recall = tpc / (nt[ci] + eps)
        precision = tpc / (tpc + fpc)
        r_curve[ci] = np.interp(-x, -conf[sel], recall[:, 0], left=0)
        p_curve[ci] = np.interp(-x, -conf[sel], precision[:, 0], left=1)
        for j in range(tp.shape[1]):
            ap[ci, j] = compute_ap(recall[:, j], precision[:, j])[0]
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    i = smooth(f1_curve.mean(0), 0.1).argmax() if n else 0
    return p_curve[:, i], r_curve[:, i], f1_curve[:, i], ap, classes.astype(int)


class DetMetrics:
    """Box metrics of one validation run (metrics.py:626-760 `Metric` + :799-906 `DetMetrics`, numbers only)."""

    keys = ["metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)"]

    def __init__(self, names=None):
        self.names = names or {}
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index = [], [], [], np.zeros((0, 10)), []
        self.speed = {"preprocess": 0.0, "inference": 0.0, "loss": 0.0, "postprocess": 0.0}

    def process(self, tp, conf, pred_cls, target_cls):
        self.p, self.r, self.f1, self.all_ap, self.ap_class_index = ap_per_class(tp, conf, pred_cls, target_cls)

    @property
    def ap50(self):
        return self.all_ap[:, 0] if len(self.all_ap) else []

    @property
    def ap(self):
        return self.all_ap.mean(1) if len(self.all_ap) else []

    @property
    def mp(self):
        return float(np.mean(self.p)) if len(self.p) else 0.0

    @property
    def mr(self):
        return float(np.mean(self.r)) if len(self.r) else 0.0

    @property
    def map50(self):
        return float(self.all_ap[:, 0].mean()) if len(self.all_ap) else 0.0

    @property
    def map75(self):
        return float(self.all_ap[:, 5].mean()) if len(self.all_ap) else 0.0

    @property
    def map(self):
        return float(self.all_ap.mean()) if len(self.all_ap) else 0.0

    def mean_results(self):
        return [self.mp, self.mr, self.map50, self.map]

    def class_result(self, i):
        return self.p[i], self.r[i], self.ap50[i], self.ap[i]

    @property
    def fitness(self):
        return float((np.array(self.mean_results()) * np.array([0.0, 0.0, 0.1, 0.9])).sum())

    @property
    def results_dict(self):
        return dict(zip(self.keys + ["fitness"], self.mean_results() + [self.fitness]))
