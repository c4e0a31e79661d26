"""numpy restatement of ops.non_max_suppression (reference ultralytics/utils/ops.py:181-332) and of
torchvision.ops.nms (torchvision 0.26.0 CPU kernel, call site ops.py:312).  TEST ORACLE.

All arithmetic is float32 in the reference's operation order.  Two deliberate choices where the reference is
ambiguous: the `max_nms` pre-sort is STABLE here (the reference's argsort is not), and ties in
torchvision's sort keep the lower index first (its sort is stable).
"""
import numpy as np

f32 = np.float32


def xywh2xyxy(x):
    """ops.py:432-449."""
    y = np.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / f32(2)
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def nms_greedy(boxes, scores, iou_threshold):
    """torchvision nms_kernel_impl (csrc/ops/cpu/nms_kernel.cpp): stable descending sort by score, then greedy
    suppression with `inter / (area_i + area_j - inter) > iou_threshold`; the float quotient is compared against
    the DOUBLE threshold.  Returns kept indices into `boxes` in score order (int64)."""
    boxes = np.asarray(boxes, dtype=f32)
    scores = np.asarray(scores, dtype=f32)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = boxes[:, 0], boxes[:, 1], boxes[:, 2], boxes[:, 3]
    areas = (x2 - x1) * (y2 - y1)
    order = np.argsort(-scores, kind="stable")
    sx1, sy1, sx2, sy2, sa = x1[order], y1[order], x2[order], y2[order], areas[order]
    suppressed = np.zeros(n, dtype=bool)
    keep = []
    thr = float(iou_threshold)
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(sx1[i], sx1[i + 1:])
        yy1 = np.maximum(sy1[i], sy1[i + 1:])
        xx2 = np.minimum(sx2[i], sx2[i + 1:])
        yy2 = np.minimum(sy2[i], sy2[i + 1:])
        w = np.maximum(f32(0), xx2 - xx1)
        h = np.maximum(f32(0), yy2 - yy1)
        inter = w * h
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = inter / (sa[i] + sa[i + 1:] - inter)
        suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        max_det=300, nc=0, max_nms=30000, max_wh=7680, nms_fn=None, return_kept=False, labels=()):
    """prediction: (B, 4+nc, A) float32.  Returns a list of (k_i, 6) float32 arrays [x1,y1,x2,y2,conf,cls]
    (and, with return_kept, the list of kept index arrays into the candidate list `x`, as torchvision returns)."""
    assert 0 <= conf_thres <= 1 and 0 <= iou_thres <= 1
    nms_fn = nms_fn or nms_greedy
    prediction = np.asarray(prediction, dtype=f32)
    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    mi = 4 + nc
    conf32 = f32(conf_thres)                                  # torch compares a float32 tensor with the scalar in fp32
    xc = prediction[:, 4:mi].max(axis=1) > conf32             # ops.py:250
    multi_label = bool(multi_label) and nc > 1                # :255
    pred = np.transpose(prediction, (0, 2, 1)).copy()         # :257
    pred[..., :4] = xywh2xyxy(pred[..., :4])                  # :259-260
    output, kept_all = [], []
    for xi in range(bs):                                      # :266
        x = pred[xi][xc[xi]]                                  # :269
        if labels and len(labels[xi]):                        # a-priori labels (autolabelling), :272-277
            lb = np.asarray(labels[xi], dtype=f32).reshape(-1, 5)
            v = np.zeros((lb.shape[0], mi), f32)
            v[:, :4] = xywh2xyxy(lb[:, 1:5])
            v[np.arange(lb.shape[0]), lb[:, 0].astype(np.int64) + 4] = 1.0
            x = np.concatenate((x, v), 0)
        if not x.shape[0]:
            output.append(np.zeros((0, 6), f32)); kept_all.append(np.zeros((0,), np.int64)); continue
        box, cls = x[:, :4], x[:, 4:mi]
        if multi_label:                                       # :285-288
            i, j = np.nonzero(cls > conf32)
            x = np.concatenate((box[i], x[i, 4 + j, None], j[:, None].astype(f32)), 1)
        else:                                                 # :289-291 (first maximum wins ties)
            j = cls.argmax(1)
            conf = cls[np.arange(cls.shape[0]), j]
            x = np.concatenate((box, conf[:, None], j[:, None].astype(f32)), 1)[conf > conf32]
        if classes is not None:                               # :294-295
            x = x[np.isin(x[:, 5], np.asarray(classes, dtype=f32))]
        n = x.shape[0]
        if not n:
            output.append(np.zeros((0, 6), f32)); kept_all.append(np.zeros((0,), np.int64)); continue
        if n > max_nms:                                       # :301-302 (stable here; see module docstring)
            x = x[np.argsort(-x[:, 4], kind="stable")[:max_nms]]
        c = x[:, 5:6] * f32(0 if agnostic else max_wh)        # :305
        boxes = x[:, :4] + c                                  # :311
        i = nms_fn(boxes, x[:, 4], iou_thres)[:max_det]       # :312-313
        output.append(x[i].astype(f32))                       # :327
        kept_all.append(np.asarray(i, dtype=np.int64))
    return (output, kept_all) if return_kept else output
