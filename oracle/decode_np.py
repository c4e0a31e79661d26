"""numpy restatement of Detect._inference (reference ultralytics/nn/modules/head.py:100-131).  TEST ORACLE."""
import numpy as np

REG_MAX = 16


def make_anchors(shapes, strides, offset=0.5):
    """reference utils/tal.py:333-345: per level (x+0.5, y+0.5) on an (h,w) grid, row-major; level-major concat."""
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = np.arange(w, dtype=np.float32) + np.float32(offset)
        sy = np.arange(h, dtype=np.float32) + np.float32(offset)
        yy, xx = np.meshgrid(sy, sx, indexing="ij")
        pts.append(np.stack((xx, yy), -1).reshape(-1, 2))
        st.append(np.full((h * w, 1), s, dtype=np.float32))
    return np.concatenate(pts), np.concatenate(st)


def decode(raw_maps, strides, nc):
    """raw_maps: list of (B, 64+nc, H_l, W_l) float arrays -> y (B, 4+nc, A) float32.

    head.py:104 view+cat; block.py:73-76 DFL: softmax over the 16 bins of each side, expectation with weights
    0..15; tal.py:348-357 dist2bbox(xywh=True): x1y1 = a - lt, x2y2 = a + rb, (c, wh); head.py:129 `* strides`;
    head.py:131 sigmoid of the class logits."""
    B = raw_maps[0].shape[0]
    no = 4 * REG_MAX + nc
    x_cat = np.concatenate([np.asarray(m, dtype=np.float32).reshape(B, no, -1) for m in raw_maps], axis=2)
    box, cls = x_cat[:, : 4 * REG_MAX], x_cat[:, 4 * REG_MAX:]
    A = x_cat.shape[2]
    anchors, stride_t = make_anchors([m.shape[2:] for m in raw_maps], strides)
    b = box.reshape(B, 4, REG_MAX, A)
    b = b - b.max(axis=2, keepdims=True)
    e = np.exp(b, dtype=np.float32)
    p = e / e.sum(axis=2, keepdims=True, dtype=np.float32)
    w = np.arange(REG_MAX, dtype=np.float32).reshape(1, 1, REG_MAX, 1)
    dist = (p * w).sum(axis=2, dtype=np.float32)                      # (B, 4, A): l, t, r, b
    a = anchors.T[None]                                               # (1, 2, A)
    x1y1 = a - dist[:, :2]
    x2y2 = a + dist[:, 2:]
    dbox = np.concatenate(((x1y1 + x2y2) / np.float32(2), x2y2 - x1y1), axis=1) * stride_t.T[None]
    prob = (np.float32(1) / (np.float32(1) + np.exp(-cls, dtype=np.float32))).astype(np.float32)
    return np.concatenate((dbox.astype(np.float32), prob), axis=1)
