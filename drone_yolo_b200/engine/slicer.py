"""InferenceSlicer: tiled detection of one large frame as ONE batch (batch dispatch of tiles).

The reference's application script (mix6.py:79-89) wraps `detection_model(tile, iou, conf, classes)` in
`supervision.InferenceSlicer(callback, slice_wh=(2160, 2160), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.7,
thread_workers=1)`: every tile of a 4K frame is a separate predict call (host letterbox, H2D, ~300 launches, D2H), then the
tile detections are moved to frame coordinates and filtered by a category-aware NMS on the host.

Here the frame crosses PCIe once; each tile is letterboxed by `dy_letterbox_u8` straight from the resident frame (a pitched
view, no host slicing) into one image of the engine's uint8 input batch; ONE engine step runs conv stack + decode + NMS for
all tiles; the per-tile rows (<= max_det each) are rescaled on the host exactly as `predict` does, moved by the tile origin
(float64, as numpy promotes them in the reference's caller) and merged by `dy_box_nms_f64` on the GPU.

Tile geometry: tiles of one frame differ in shape at the right / bottom edges, so — exactly like the reference's predictor
given a LIST of differently shaped images (predictor.py:157-163: `auto` only when all shapes are equal) — every tile is
letterboxed onto the full imgsz x imgsz canvas.  (mix6.py's one-tile-per-call form pads each tile only to a stride
multiple; detections near the padded border can differ in the last bits of their scores.)

The constructor mirrors supervision's (slice_wh, overlap_ratio_wh, iou_threshold, thread_workers accepted and ignored); the
first argument is the YOLO model instead of a per-tile callback, and `__call__(image)` returns ONE `Results` in frame
coordinates instead of `sv.Detections` (`sv.Detections.from_ultralytics(result)` reads `.boxes.xyxy/.conf/.cls`).
"""
from __future__ import annotations

import time

import numpy as np
import torch

from .. import _C
from .. import kernels as K
from ..utils import ops
from .predictor import DetectionPredictor, letterbox_geometry
from .results import Results


def generate_offsets(resolution_wh, slice_wh, overlap_ratio_wh) -> np.ndarray:
    """(T, 4) int64 [x_min, y_min, x_max, y_max], row-major over the tile grid: stride = slice - int(ratio * slice), origins
    every stride while inside the frame, far edges clipped (supervision InferenceSlicer._generate_offset)."""
    sw, sh = int(slice_wh[0]), int(slice_wh[1])
    iw, ih = int(resolution_wh[0]), int(resolution_wh[1])
    if sw <= 0 or sh <= 0:
        raise ValueError(f"slice_wh {slice_wh} must be positive")
    if not (0 <= overlap_ratio_wh[0] < 1 and 0 <= overlap_ratio_wh[1] < 1):
        raise ValueError(f"overlap_ratio_wh {overlap_ratio_wh} must be in [0, 1)")
    stride_w = sw - int(overlap_ratio_wh[0] * sw)
    stride_h = sh - int(overlap_ratio_wh[1] * sh)
    out = [(x, y, min(x + sw, iw), min(y + sh, ih)) for y in range(0, ih, stride_h) for x in range(0, iw, stride_w)]
    return np.asarray(out, dtype=np.int64).reshape(-1, 4)


class InferenceSlicer:
    def __init__(self, model, slice_wh=(320, 320), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.5, class_agnostic=False,
                 thread_workers=1, **predict_kwargs):
        self.model = model
        self.slice_wh = (int(slice_wh[0]), int(slice_wh[1]))
        self.overlap_ratio_wh = (float(overlap_ratio_wh[0]), float(overlap_ratio_wh[1]))
        if not 0.0 <= float(iou_threshold) <= 1.0:
            raise AssertionError(f"Invalid IoU {iou_threshold}, valid values are between 0.0 and 1.0")
        self.iou_threshold = float(iou_threshold)
        self.class_agnostic = bool(class_agnostic)
        self.predict_kwargs = dict(predict_kwargs)
        self._frame: dict = {}

    # the predictor of the wrapped YOLO object, created / updated exactly as Model.predict does (engine/model.py:501-560)
    def _predictor(self) -> DetectionPredictor:
        m = self.model
        args = {**m.overrides, "conf": 0.25, "batch": 1, "mode": "predict", **self.predict_kwargs}
        args.pop("model", None)
        if m.predictor is None:
            m.predictor = DetectionPredictor(overrides=args)
            m.predictor.setup_model(model=m.model)
        else:
            m.predictor.update_args(args)
        return m.predictor

    MAX_MERGE_ROWS = 16384                     # dy_box_nms_f64: one CTA scans the `removed` set, 256 words of 64 rows

    def _merge_keep(self, rows: np.ndarray, device) -> np.ndarray:
        """Keep mask of the merge NMS over frame rows (n, 6) float64.  More rows than one launch takes (many small tiles at a
        low confidence threshold) are merged category by category — categories never interact unless class_agnostic."""
        n = len(rows)
        if n <= self.MAX_MERGE_ROWS:
            return K.box_nms_f64(torch.from_numpy(rows).to(device), self.iou_threshold, self.class_agnostic).cpu().numpy()
        if self.class_agnostic:
            raise _C.DroneYoloError(f"InferenceSlicer: {n} rows to merge class-agnostically, at most {self.MAX_MERGE_ROWS} per "
                                    "launch: raise conf, lower max_det or use larger tiles")
        keep = np.zeros(n, dtype=bool)
        for c in np.unique(rows[:, 5]):
            idx = np.nonzero(rows[:, 5] == c)[0]
            if len(idx) > self.MAX_MERGE_ROWS:
                raise _C.DroneYoloError(f"InferenceSlicer: {len(idx)} rows of class {int(c)} to merge, at most "
                                        f"{self.MAX_MERGE_ROWS} per launch: raise conf, lower max_det or use larger tiles")
            sub = np.ascontiguousarray(rows[idx])
            keep[idx] = K.box_nms_f64(torch.from_numpy(sub).to(device), self.iou_threshold, False).cpu().numpy()
        return keep

    def _upload(self, image: np.ndarray, device) -> torch.Tensor:
        key = image.shape
        if key not in self._frame:
            self._frame.clear()
            self._frame[key] = (torch.empty(key, dtype=torch.uint8).pin_memory(),
                                torch.empty(key, dtype=torch.uint8, device=device))
        host, dev = self._frame[key]
        src = torch.from_numpy(np.ascontiguousarray(image))
        rows = max(1, (4 << 20) // max(1, image.shape[1] * 3))     # ~4 MB bands: band k crosses PCIe while band k+1 is staged
        for r0 in range(0, image.shape[0], rows):
            host[r0:r0 + rows].copy_(src[r0:r0 + rows])
            dev[r0:r0 + rows].copy_(host[r0:r0 + rows], non_blocking=True)
        return dev

    def __call__(self, image: np.ndarray) -> Results:
        if not isinstance(image, np.ndarray) or image.ndim != 3 or image.shape[2] != 3 or image.dtype != np.uint8:
            raise _C.DroneYoloError("InferenceSlicer: the frame must be a uint8 (h, w, 3) BGR array")
        p = self._predictor()
        a = p.args
        t0 = time.perf_counter()
        offsets = generate_offsets((image.shape[1], image.shape[0]), self.slice_wh, self.overlap_ratio_wh)
        T = len(offsets)
        shape = (a.imgsz, a.imgsz) if isinstance(a.imgsz, int) else tuple(a.imgsz)
        tile_shapes = [(int(y1 - y0), int(x1 - x0)) for x0, y0, x1, y1 in offsets]
        same = len(set(tile_shapes)) == 1
        geo = [letterbox_geometry(s, shape, auto=same) for s in tile_shapes]
        H, W = geo[0][4], geo[0][5]
        with p._lock, torch.inference_mode():
            frame = self._upload(image, p.device)
            p._flush()                                 # nothing of a stream_inference loop is in flight on this engine's slots
            eng = p.engine_for(T, H, W, torch.uint8)
            for i, ((x0, y0, x1, y1), g) in enumerate(zip(offsets.tolist(), geo)):
                K.letterbox_u8(frame[y0:y1, x0:x1], eng.image_slots[0][i], g[0], g[1], g[2], g[3])
            # tile-coordinate rows straight from the NMS output phase (scale_boxes + clip_boxes per tile, dy_nms_desc.rescale)
            eng.rescale_slots[0].copy_(ops.rescale_params((H, W), tile_shapes))
            t1 = time.perf_counter()
            out, counts = p.inference(eng, 0)
            host, n = out.cpu(), counts.cpu().tolist()
            per_tile = [host[i, :n[i]] for i in range(T)]
            t2 = time.perf_counter()
            moved = []
            for r, (x0, y0, _, _) in zip(per_tile, offsets.tolist()):
                rows = np.asarray(r, dtype=np.float64)
                if len(rows):
                    rows = rows.copy()
                    rows[:, [0, 2]] += float(x0)
                    rows[:, [1, 3]] += float(y0)
                    moved.append(rows)
            rows = np.concatenate(moved, 0) if moved else np.zeros((0, 6))
            if len(rows):
                rows = rows[self._merge_keep(rows, p.device)]
            t3 = time.perf_counter()
        res = Results(image, path="frame.jpg", names=self.model.model.names, boxes=torch.from_numpy(rows),
                      orig_shape=image.shape[:2])
        res.speed = {"preprocess": (t1 - t0) * 1e3, "inference": (t2 - t1) * 1e3, "postprocess": (t3 - t2) * 1e3}
        res.tiles = offsets
        return res
