"""numpy restatement of the tiled-frame dispatch the reference's application script drives (mix6.py:79-89):
`supervision.InferenceSlicer(callback, slice_wh, overlap_ratio_wh, iou_threshold)`.  TEST ORACLE.

`supervision` is a third-party dependency that is NOT under /root/reference and is not installed in this image
(imported at mix6.py:2, unpinned): its published algorithm (InferenceSlicer._generate_offset, move_boxes,
Detections.merge, Detections.with_nms -> box_non_max_suppression, box_iou_batch) is restated here from its
documentation — PARITY UNPINNED for this step; the anchors are the call site's arguments (mix6.py:84-89: slice_wh
(2160, 2160), overlap (0.2, 0.2), iou 0.7), hand-computed known answers, and — for the greedy rule itself — agreement with the
installed torchvision.ops.nms on tie-free float64 rows (tests/test_oracle_golden.py).

Precision: tile boxes are float32 (Results.boxes.xyxy); adding the integer tile origin promotes them to float64, so the
merge NMS runs in float64.  Ranking: `np.flip(argsort(conf))` — with a stable ascending sort, equal scores rank the
HIGHER row first; that is the rule fixed here (numpy's default sort is not stable for more than 16 rows).
"""
import numpy as np


def generate_offsets(resolution_wh, slice_wh, overlap_ratio_wh):
    """(T, 4) int64 [x_min, y_min, x_max, y_max] of every tile, row-major over the grid (InferenceSlicer._generate_offset):
    stride = slice - int(overlap_ratio * slice); origins arange(0, size, stride); far edges clipped to the frame."""
    sw, sh = int(slice_wh[0]), int(slice_wh[1])
    iw, ih = int(resolution_wh[0]), int(resolution_wh[1])
    ow, oh = int(overlap_ratio_wh[0] * sw), int(overlap_ratio_wh[1] * sh)
    ws = np.arange(0, iw, sw - ow)
    hs = np.arange(0, ih, sh - oh)
    xmin, ymin = np.meshgrid(ws, hs)
    xmax = np.clip(xmin + sw, 0, iw)
    ymax = np.clip(ymin + sh, 0, ih)
    return np.stack([xmin, ymin, xmax, ymax], axis=-1).reshape(-1, 4).astype(np.int64)


def move_rows(rows, offset_xy):
    """Tile rows (k, 6) float32 -> frame rows (k, 6) float64: xyxy + [x0, y0, x0, y0] (move_boxes)."""
    out = np.asarray(rows).astype(np.float64)
    out[:, [0, 2]] += float(offset_xy[0])
    out[:, [1, 3]] += float(offset_xy[1])
    return out


def box_iou_batch(a, b):
    """(n, m) IoU: inter / (area_a[:, None] + area_b - inter) with inter = prod(clip(min(br) - max(tl), 0))."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    tl = np.maximum(a[:, None, :2], b[None, :, :2])
    br = np.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = np.clip(br - tl, 0, None)
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return inter / (area_a[:, None] + area_b[None, :] - inter)


def box_nms_keep(rows, iou_threshold, class_agnostic=False):
    """Keep mask (n,) bool in the ORIGINAL row order (box_non_max_suppression): rank by conf descending (equal conf:
    higher row first), a kept row removes every other row of its category with IoU > threshold (strict; NaN is False)."""
    rows = np.asarray(rows, dtype=np.float64).reshape(-1, 6)
    n = rows.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=bool)
    sort_index = np.flip(rows[:, 4].argsort(kind="stable"))
    r = rows[sort_index]
    cats = np.zeros(n) if class_agnostic else r[:, 5]
    ious = box_iou_batch(r[:, :4], r[:, :4])
    ious[np.arange(n), np.arange(n)] = 0.0                      # `ious - eye`: a row never removes itself
    keep = np.ones(n, dtype=bool)
    for i in range(n):
        if not keep[i]:
            continue
        keep &= ~((ious[i] > iou_threshold) & (cats == cats[i]))
    return keep[sort_index.argsort(kind="stable")]


def merge_tiles(tile_rows, offsets, iou_threshold, class_agnostic=False):
    """Per-tile (k_t, 6) rows + (T, 4) offsets -> merged (m, 6) float64 frame rows, tile-major order (Detections.merge
    keeps the input order, with_nms filters it)."""
    moved = [move_rows(r, o[:2]) for r, o in zip(tile_rows, offsets) if len(r)]
    if not moved:
        return np.zeros((0, 6))
    rows = np.concatenate(moved, 0)
    return rows[box_nms_keep(rows, iou_threshold, class_agnostic)]
