"""Post-processing ops with the reference's signatures (ultralytics/utils/ops.py).

`non_max_suppression` keeps the reference's argument list, output layout and side effects, but the whole batch is
processed by two CUDA launches (dy_nms) instead of a Python loop over images around torchvision.ops.nms.
The small box helpers (`scale_boxes`, `clip_boxes`, `xywh2xyxy`) act on the <= max_det rows per image that leave NMS;
they are host-side bookkeeping, not kernels.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from .. import kernels as K
from .._C import DroneYoloError

_nms_cache: dict = {}


def make_divisible(x, divisor):
    if isinstance(divisor, torch.Tensor):
        divisor = int(divisor.max())
    return math.ceil(x / divisor) * divisor


def xywh2xyxy(x):
    """(cx, cy, w, h) -> (x1, y1, x2, y2) (reference ops.py:432-449)."""
    assert x.shape[-1] == 4, f"input shape last dimension expected 4 but input shape is {x.shape}"
    y = torch.empty_like(x) if isinstance(x, torch.Tensor) else np.empty_like(x)
    xy = x[..., :2]
    wh = x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def clip_boxes(boxes, shape):
    """Clamp xyxy boxes to an image of `shape` (h, w) (reference ops.py:335-354)."""
    if isinstance(boxes, torch.Tensor):
        boxes[..., 0] = boxes[..., 0].clamp(0, shape[1])
        boxes[..., 1] = boxes[..., 1].clamp(0, shape[0])
        boxes[..., 2] = boxes[..., 2].clamp(0, shape[1])
        boxes[..., 3] = boxes[..., 3].clamp(0, shape[0])
    else:
        boxes[..., [0, 2]] = boxes[..., [0, 2]].clip(0, shape[1])
        boxes[..., [1, 3]] = boxes[..., [1, 3]].clip(0, shape[0])
    return boxes


def scale_boxes(img1_shape, boxes, img0_shape, ratio_pad=None, padding=True, xywh=False):
    """Rescale boxes from the network input shape to the original image shape (reference ops.py:92-127)."""
    if ratio_pad is None:
        gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
        pad = (
            round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1),
            round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1),
        )
    else:
        gain = ratio_pad[0][0]
        pad = ratio_pad[1]
    if padding:
        boxes[..., 0] -= pad[0]
        boxes[..., 1] -= pad[1]
        if not xywh:
            boxes[..., 2] -= pad[0]
            boxes[..., 3] -= pad[1]
    boxes[..., :4] /= gain
    return clip_boxes(boxes, img0_shape)


def scale_boxes_batch(img1_shape, boxes, img0_shapes):
    """`scale_boxes` for a whole batch at once: boxes (B, k, 4) float32 xyxy (CPU tensor, modified in place), img0_shapes a list
    of B (h, w).  Element for element the arithmetic of the per-image call (reference ops.py:92-127: `-= pad`, `/= gain` with the
    Python scalar taken as float32, clamp to the image), in three broadcast tensor ops instead of ~12 small ones per image."""
    shapes = [tuple(s[:2]) for s in img0_shapes]
    uniq = {}
    for s in shapes:
        if s not in uniq:
            g = min(img1_shape[0] / s[0], img1_shape[1] / s[1])
            px = round((img1_shape[1] - s[1] * g) / 2 - 0.1)
            py = round((img1_shape[0] - s[0] * g) / 2 - 0.1)
            uniq[s] = ([px, py, px, py], g, [s[1], s[0], s[1], s[0]])
    if len(uniq) == 1:
        pad, g, hi = next(iter(uniq.values()))
        pad_t = torch.tensor(pad, dtype=torch.float32)
        g_t = torch.tensor(g, dtype=torch.float64).to(torch.float32)
        hi_t = torch.tensor(hi, dtype=torch.float32)
    else:
        pad_t = torch.tensor([uniq[s][0] for s in shapes], dtype=torch.float32)[:, None, :]
        g_t = torch.tensor([uniq[s][1] for s in shapes], dtype=torch.float64).to(torch.float32)[:, None, None]
        hi_t = torch.tensor([uniq[s][2] for s in shapes], dtype=torch.float32)[:, None, :]
    boxes -= pad_t
    boxes /= g_t
    torch.minimum(boxes, hi_t, out=boxes)
    torch.maximum(boxes, torch.zeros((), dtype=torch.float32), out=boxes)
    return boxes


def rescale_params(img1_shape, img0_shapes) -> torch.Tensor:
    """(B, 8) float32 rows (pad_x, pad_y, gain, w0, h0, 0, 0, 0): what `scale_boxes` + `clip_boxes` (reference ops.py:92-127,
    335-354) need per image, for the fused post-step of dy_nms (`dy_nms_desc.rescale`).  One row per DISTINCT shape is computed
    with the reference's Python arithmetic (float64 gain, `round(... - 0.1)`), then broadcast."""
    out = np.zeros((len(img0_shapes), 8), dtype=np.float32)
    rows = {}
    for i, s in enumerate(img0_shapes):
        s = (s[0], s[1])
        r = rows.get(s)
        if r is None:
            g = min(img1_shape[0] / s[0], img1_shape[1] / s[1])
            r = rows[s] = (round((img1_shape[1] - s[1] * g) / 2 - 0.1), round((img1_shape[0] - s[0] * g) / 2 - 0.1), g, s[1], s[0])
        out[i, :5] = r
    return torch.from_numpy(out)


def convert_torch2numpy_batch(batch: torch.Tensor) -> np.ndarray:
    """(B,C,H,W) float 0..1 -> (B,H,W,C) uint8 (reference ops.py:841-851)."""
    return (batch.permute(0, 2, 3, 1).contiguous() * 255).clamp(0, 255).to(torch.uint8).cpu().numpy()


def nms_padded(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
               max_det=300, nc=0, max_nms=30000, max_wh=7680, in_place=False):
    """The kernel-level call: returns padded device tensors (out (B,max_det,6), counts (B,), kept (B,max_det))
    without any host synchronisation.  Used by the predictor; `non_max_suppression` wraps it."""
    B, ch, A = prediction.shape
    nc = nc or (ch - 4)
    if ch - nc - 4 != 0:
        raise DroneYoloError("non_max_suppression: mask channels (nm > 0) are outside the Drone-YOLO detect path")
    ml = bool(multi_label) and nc > 1
    key = (B, nc, A, int(max_det), ml, str(prediction.device))
    bufs = _nms_cache.get(key)
    if bufs is None:
        if len(_nms_cache) > 8:
            _nms_cache.clear()
        bufs = _nms_cache[key] = K.NmsBuffers(B, nc, A, int(max_det), ml, prediction.device)
    return K.nms(prediction, conf_thres, iou_thres, max_det=max_det, max_nms=max_nms, max_wh=max_wh, agnostic=agnostic,
                 multi_label=ml, classes=classes, in_place=in_place, bufs=bufs)


def _append_labels(pred: torch.Tensor, labels, nc: int) -> torch.Tensor:
    """(B, 4+nc, A) -> (B, 4+nc, A + L): label rows (cls, cx, cy, w, h) of image i as columns A.. of image i."""
    B, ch, A = pred.shape
    if len(labels) != B:
        raise DroneYoloError(f"non_max_suppression: {len(labels)} label lists for a batch of {B}")
    L = max(len(lb) for lb in labels)
    aug = torch.zeros((B, ch, A + (L + 3) // 4 * 4), device=pred.device, dtype=torch.float32)
    aug[:, :, :A] = pred
    for i, lb in enumerate(labels):
        if not len(lb):
            continue
        lb = torch.as_tensor(lb, dtype=torch.float32, device=pred.device).reshape(-1, 5)
        cls = lb[:, 0].long()
        if int(cls.min()) < 0 or int(cls.max()) >= nc:
            raise DroneYoloError("non_max_suppression: label class outside [0, nc)")
        cols = A + torch.arange(lb.shape[0], device=pred.device)
        aug[i, :4, cols] = lb[:, 1:5].t()
        aug[i, 4 + cls, cols] = 1.0
    return aug


def non_max_suppression(
    prediction,
    conf_thres=0.25,
    iou_thres=0.45,
    classes=None,
    agnostic=False,
    multi_label=False,
    labels=(),
    max_det=300,
    nc=0,
    max_time_img=0.05,
    max_nms=30000,
    max_wh=7680,
    in_place=True,
    rotated=False,
    end2end=False,
):
    """Batched NMS with the reference's signature and output (reference ops.py:181-332).

    Returns a list of B tensors (k_i, 6): x1, y1, x2, y2, confidence, class; score-descending, equal scores in
    candidate order.  `max_time_img` is accepted and ignored: the kernels have no wall-clock cut-off, so no image
    is ever silently dropped (reference ops.py:328-330).
    """
    assert 0 <= conf_thres <= 1, f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0"
    assert 0 <= iou_thres <= 1, f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0"
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]
    if rotated or end2end or prediction.shape[-1] == 6:
        raise DroneYoloError("non_max_suppression: rotated / end2end inputs are outside the Drone-YOLO detect path")
    if not prediction.is_cuda:
        raise DroneYoloError("non_max_suppression runs on CUDA tensors only; there is no CPU fallback")

    work = prediction
    if work.dtype != torch.float32 or not work.is_contiguous():
        work = prediction.float().contiguous()
    if labels and any(len(lb) for lb in labels):
        # A-priori labels (autolabelling, the validator's `save_hybrid` path; reference ops.py:272-277): every label row
        # (cls, cx, cy, w, h) joins its image's candidates behind the anchors with score 1.0 for its class.  Here they
        # become extra anchor columns of a widened copy (unused columns score 0 and never pass `conf`), so candidate
        # order, tie order and kept indices are the reference's.
        work = _append_labels(work, labels, nc or work.shape[1] - 4)
        out, counts, _ = nms_padded(work, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc, max_nms,
                                    max_wh, in_place=False)
        if in_place:                              # the caller's tensor still turns xywh -> xyxy (:259-260)
            xy, wh = prediction[:, :2].float(), prediction[:, 2:4].float() / 2
            prediction[:, :2], prediction[:, 2:4] = (xy - wh).to(prediction.dtype), (xy + wh).to(prediction.dtype)
    else:
        out, counts, _ = nms_padded(work, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc, max_nms,
                                    max_wh, in_place=in_place)
        if in_place and work is not prediction:   # keep the reference's side effect on the caller's tensor (:259-260)
            prediction[:, :4] = work[:, :4].to(prediction.dtype)
    n = counts.tolist()                       # the one host sync: per-image row counts
    out = out.clone()
    return [out[i, : n[i]] for i in range(out.shape[0])]
