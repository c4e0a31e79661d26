"""The same-box GPU bar (BASELINE.md 4 "Secondary GPU bar"): what the reference's PyTorch branch executes
on a CUDA device - torch eager, channels_last, cuDNN convolutions, Detect decode in aten ops, ops.non_max_suppression
with torchvision.ops.nms per image.  TEST / BENCH INFRASTRUCTURE ("port": the reference tree does not travel to the GPU
box, so its module forwards are restated by oracle/torch_ref.py and its decode / NMS here); never imported by the product.

Cited reference code:
  module forwards                  oracle/torch_ref.py (conv.py:49-55, block.py:187-350,1480-1490, tasks.py:134-161)
  Detect._inference / decode       ultralytics/nn/modules/head.py:100-131, block.py:73-76 (DFL), utils/tal.py:333-357
  ops.non_max_suppression          ultralytics/utils/ops.py:181-332 (torchvision.ops.nms call site :312)
  AutoBackend fuse / half          ultralytics/nn/autobackend.py:149-159
"""
import torch

from . import torch_ref


def make_anchors(feats, strides, offset=0.5):
    """utils/tal.py:333-345."""
    pts, st = [], []
    for f, s in zip(feats, strides):
        h, w = f.shape[2:]
        sx = torch.arange(w, device=f.device, dtype=f.dtype) + offset
        sy = torch.arange(h, device=f.device, dtype=f.dtype) + offset
        sy, sx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), s, dtype=f.dtype, device=f.device))
    return torch.cat(pts), torch.cat(st)


def decode(raw, strides, nc, reg_max=16):
    """head.py:100-131: cat the levels, DFL softmax expectation (block.py:73-76), dist2bbox(xywh) * stride, class sigmoid."""
    B = raw[0].shape[0]
    no = 4 * reg_max + nc
    x_cat = torch.cat([r.reshape(B, no, -1) for r in raw], 2)
    box, cls = x_cat.split((4 * reg_max, nc), 1)
    anchors, st = make_anchors(raw, strides)
    anchors, st = anchors.transpose(0, 1).unsqueeze(0), st.transpose(0, 1)
    A = box.shape[2]
    w = torch.arange(reg_max, dtype=box.dtype, device=box.device).view(1, reg_max, 1, 1)
    dist = (box.view(B, 4, reg_max, A).transpose(2, 1).softmax(1) * w).sum(1)            # the DFL 1x1 conv with weights 0..15
    lt, rb = dist.chunk(2, 1)
    x1y1, x2y2 = anchors - lt, anchors + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


def xywh2xyxy(x):
    y = torch.empty_like(x)
    xy, wh = x[..., :2], x[..., 2:] / 2
    y[..., :2] = xy - wh
    y[..., 2:] = xy + wh
    return y


def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False, multi_label=False,
                        max_det=300, nc=0, max_nms=30000, max_wh=7680):
    """ops.py:181-332 on the device the prediction lives on, torchvision.ops.nms per image (:312)."""
    import torchvision

    bs = prediction.shape[0]
    nc = nc or (prediction.shape[1] - 4)
    mi = 4 + nc
    xc = prediction[:, 4:mi].amax(1) > conf_thres
    multi_label &= nc > 1
    prediction = prediction.transpose(-1, -2)
    prediction = torch.cat((xywh2xyxy(prediction[..., :4]), prediction[..., 4:]), dim=-1)
    output = [torch.zeros((0, 6), device=prediction.device)] * bs
    for xi, x in enumerate(prediction):
        x = x[xc[xi]]
        if not x.shape[0]:
            continue
        box, cls = x[:, :4], x[:, 4:mi]
        if multi_label:
            i, j = torch.where(cls > conf_thres)
            x = torch.cat((box[i], x[i, 4 + j, None], j[:, None].float()), 1)
        else:
            conf, j = cls.max(1, keepdim=True)
            x = torch.cat((box, conf, j.float()), 1)[conf.view(-1) > conf_thres]
        if classes is not None:
            x = x[(x[:, 5:6] == torch.tensor(classes, device=x.device)).any(1)]
        n = x.shape[0]
        if not n:
            continue
        if n > max_nms:
            x = x[x[:, 4].argsort(descending=True)[:max_nms]]
        c = x[:, 5:6] * (0 if agnostic else max_wh)
        i = torchvision.ops.nms(x[:, :4] + c, x[:, 4], iou_thres)[:max_det]
        output[xi] = x[i]
    return output


@torch.no_grad()
def forward(model, x):
    """The module walk of torch_ref.forward with the decode kept on the device: x (B,3,H,W) in the model's dtype ->
    y (B, 4+nc, A) fp32 (the reference's AutoBackend returns the head output as the model computes it; NMS runs on .float())."""
    y = []
    for m in model.model:
        if m.f != -1:
            x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
        if type(m).__name__ == "Detect":
            raw = torch_ref.detect_raw(m, x)
            return decode([r.float() for r in raw], [float(s) for s in m.stride.tolist()], m.nc)
        x = torch_ref.module_forward(m, x)
        y.append(x if m.i in model.save else None)
    raise RuntimeError("model has no Detect head")


def prepare(model, device, dtype=torch.bfloat16, deploy=False):
    """AutoBackend's preparation (autobackend.py:149-159): fuse (Conv+BN folded; RepVGG blocks stay two-branch unless `deploy`,
    SURVEY.md F5), move to the device, cast, channels_last weights."""
    import copy

    m = copy.deepcopy(model).eval()
    m = m.fuse(verbose=False) if deploy else torch_ref.fuse_like_reference(m)
    return m.to(device=device, dtype=dtype).to(memory_format=torch.channels_last)


@torch.no_grad()
def step(model, images, conf, iou, max_det, dtype=torch.bfloat16):
    """One predict step on resident fp32 images: cast + channels_last, conv stack, decode, NMS per image."""
    x = images.to(dtype).contiguous(memory_format=torch.channels_last)
    y = forward(model, x)
    return non_max_suppression(y, conf, iou, max_det=max_det)
