"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's letterbox preprocess for one uint8 image.

Follows LetterBox.__call__ (ultralytics/data/augment.py:1544-1610: ratio, rounding of the new size and of the border,
cv2.resize(INTER_LINEAR), cv2.copyMakeBorder(value=114)) and BasePredictor.preprocess's BGR->RGB, HWC->CHW
(ultralytics/engine/predictor.py:127-131).  The resize is restated from OpenCV's 8-bit INTER_LINEAR (third-party:
opencv-python, requirements.txt `opencv-python>=4.6.0`, installed 4.13.0; modules/imgproc/src/resize.cpp: 11-bit
fixed-point coefficients, `((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2`).  Pinned against cv2 itself in
tests/test_host_cpu.py: bit-exact when shrinking, within one level on < 0.1 % of the pixels when enlarging.
Only tests/ may import this module.
"""
import numpy as np


def geometry(shape, new_shape, stride=32, auto=False):
    """(new_w, new_h, left, top, out_h, out_w) of LetterBox for an image of `shape` (h, w) (augment.py:1573-1598)."""
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_w, new_h = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    dw /= 2
    dh /= 2
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return new_w, new_h, left, top, new_h + top + bottom, new_w + left + right


def _coefs(n_dst, n_src):
    scale = 1.0 / (n_dst / n_src)
    f = ((np.arange(n_dst) + 0.5) * scale - 0.5).astype(np.float32)
    i = np.floor(f).astype(np.int64)
    fr = (f - i.astype(np.float32)).astype(np.float32)
    lo, hi = i < 0, i >= n_src - 1
    i[lo] = 0
    fr[lo] = 0
    i[hi] = n_src - 1
    fr[hi] = 0
    a1 = np.rint(fr * np.float32(2048)).astype(np.int64)
    a0 = np.rint((np.float32(1) - fr) * np.float32(2048)).astype(np.int64)
    return i, np.minimum(i + 1, n_src - 1), a0, a1


def resize_linear_u8(src, new_w, new_h):
    ix, ix1, ax0, ax1 = _coefs(new_w, src.shape[1])
    iy, iy1, ay0, ay1 = _coefs(new_h, src.shape[0])
    s = src.astype(np.int64)
    rows = s[:, ix, :] * ax0[None, :, None] + s[:, ix1, :] * ax1[None, :, None]
    out = (((ay0[:, None, None] * (rows[iy] >> 4)) >> 16) + ((ay1[:, None, None] * (rows[iy1] >> 4)) >> 16) + 2) >> 2
    return out.astype(np.uint8)


def letterbox_chw_rgb(im, new_shape, stride=32, auto=False, fill=114):
    """uint8 HWC BGR -> uint8 CHW RGB letterboxed canvas."""
    new_w, new_h, left, top, H, W = geometry(im.shape[:2], new_shape, stride, auto)
    canvas = np.full((H, W, 3), fill, np.uint8)
    canvas[top:top + new_h, left:left + new_w] = resize_linear_u8(im, new_w, new_h)
    return np.ascontiguousarray(canvas[..., ::-1].transpose(2, 0, 1))
