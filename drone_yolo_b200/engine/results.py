"""Result containers with the reference's accessors (ultralytics/engine/results.py:187-271, 1004-1198):
`Results(orig_img, path, names, boxes)` and `Boxes` over an (n, 6) array [x1, y1, x2, y2, conf, cls].
Plotting / saving / exporters are outside the inference hot path and not provided."""
from __future__ import annotations

import numpy as np
import torch


class Boxes:
    def __init__(self, boxes, orig_shape):
        if boxes.ndim == 1:
            boxes = boxes[None, :]
        assert boxes.shape[-1] == 6, f"expected 6 values per box but got {boxes.shape[-1]}"
        self.data = boxes
        self.orig_shape = orig_shape
        self.is_track = False

    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return None

    @property
    def xywh(self):
        b = self.xyxy
        out = b.clone() if isinstance(b, torch.Tensor) else b.copy()
        out[:, 0] = (b[:, 0] + b[:, 2]) / 2
        out[:, 1] = (b[:, 1] + b[:, 3]) / 2
        out[:, 2] = b[:, 2] - b[:, 0]
        out[:, 3] = b[:, 3] - b[:, 1]
        return out

    def _norm(self, b):
        out = b.clone() if isinstance(b, torch.Tensor) else b.copy()
        out[:, [0, 2]] /= self.orig_shape[1]
        out[:, [1, 3]] /= self.orig_shape[0]
        return out

    @property
    def xyxyn(self):
        return self._norm(self.xyxy)

    @property
    def xywhn(self):
        return self._norm(self.xywh)

    @property
    def shape(self):
        return self.data.shape

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return Boxes(self.data[idx], self.orig_shape)

    def cpu(self):
        return self if isinstance(self.data, np.ndarray) else Boxes(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self if isinstance(self.data, np.ndarray) else Boxes(self.data.cpu().numpy(), self.orig_shape)

    def cuda(self):
        return Boxes(torch.as_tensor(self.data).cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return Boxes(torch.as_tensor(self.data).to(*args, **kwargs), self.orig_shape)


_NO_SPEED = {"preprocess": None, "inference": None, "postprocess": None}


class Results:
    """One image's detections.  `orig_img` may be a callable producing the uint8 HWC image lazily (tensor sources:
    the reference pays a D2H round trip of the whole input batch for it, utils/ops.py:851)."""

    def __init__(self, orig_img, path, names, boxes=None, orig_shape=None, speed=None):
        self._orig_img = orig_img
        self.orig_shape = orig_shape if orig_shape is not None else orig_img.shape[:2]
        # (n, 6) rows, or (block (B, max_det, 6), image index, row count): the slice and the Boxes view are made on first access
        # (rank 0 of an 8-GPU run receives 512 Results per engine step; most consumers touch a few fields of each)
        self._rows = boxes
        self._boxes = None
        self.masks = self.probs = self.keypoints = self.obb = None
        self.speed = speed or _NO_SPEED
        self.names = names
        self.path = path
        self.save_dir = None

    def _resolve(self):
        r = self._rows
        if type(r) is tuple:
            block, i, cnt = r
            r = self._rows = block[i, :cnt]
        return r

    @property
    def boxes(self):
        if self._boxes is None and self._rows is not None:
            self._boxes = Boxes(self._resolve(), self.orig_shape)
        return self._boxes

    @boxes.setter
    def boxes(self, value):
        self._boxes = value
        self._rows = value.data if value is not None else None

    @property
    def orig_img(self):
        if callable(self._orig_img):
            self._orig_img = self._orig_img()
        return self._orig_img

    def __len__(self):
        r = self._rows
        if type(r) is tuple:
            return r[2]
        return 0 if r is None else (1 if r.ndim == 1 else r.shape[0])

    def _with(self, boxes):
        r = Results(self._orig_img, self.path, self.names, None, self.orig_shape, self.speed)
        r.boxes = boxes
        return r

    def cpu(self):
        return self._with(self.boxes.cpu())

    def numpy(self):
        return self._with(self.boxes.numpy())

    def to(self, *a, **k):
        return self._with(self.boxes.to(*a, **k))

    def summary(self):
        b = self.boxes.numpy()
        return [{"name": self.names[int(c)], "class": int(c), "confidence": float(p),
                 "box": {"x1": float(x[0]), "y1": float(x[1]), "x2": float(x[2]), "y2": float(x[3])}}
                for x, p, c in zip(b.xyxy, b.conf, b.cls)]

    def verbose(self):
        if len(self) == 0:
            return "(no detections), "
        cls = self.boxes.numpy().cls.astype(int)
        return "".join(f"{(cls == c).sum()} {self.names[int(c)]}{'s' * ((cls == c).sum() > 1)}, " for c in np.unique(cls))
