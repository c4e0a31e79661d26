"""Anchor helpers (reference ultralytics/utils/tal.py:333-357).

The decode kernel derives anchors from the anchor index, so these exist for API parity and host-side checks only
(tiny, shape-only arithmetic; not on the device hot path)."""
import torch


def make_anchors(feats, strides, grid_cell_offset=0.5):
    """Anchor points (A,2) as (x+0.5, y+0.5) per level, level-major, and the stride of every anchor (A,1)."""
    anchor_points, stride_tensor = [], []
    dtype, device = feats[0].dtype, feats[0].device
    for i, stride in enumerate(strides):
        h, w = feats[i].shape[2:] if isinstance(feats, (list, tuple)) else (int(feats[i][0]), int(feats[i][1]))
        sx = torch.arange(end=w, device=device, dtype=dtype) + grid_cell_offset
        sy = torch.arange(end=h, device=device, dtype=dtype) + grid_cell_offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        anchor_points.append(torch.stack((sx, sy), -1).view(-1, 2))
        stride_tensor.append(torch.full((h * w, 1), float(stride), dtype=dtype, device=device))
    return torch.cat(anchor_points), torch.cat(stride_tensor)


def dist2bbox(distance, anchor_points, xywh=True, dim=-1):
    """(l,t,r,b) distances -> boxes: x1y1 = a - lt, x2y2 = a + rb; xywh returns (centre, size)."""
    lt, rb = distance.chunk(2, dim)
    x1y1 = anchor_points - lt
    x2y2 = anchor_points + rb
    if xywh:
        return torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), dim)
    return torch.cat((x1y1, x2y2), dim)
