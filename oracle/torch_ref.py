"""Plain PyTorch fp32 (CPU) restatement of the reference's module forwards for the Drone-YOLO graphs.  TEST ORACLE
and CPU baseline ("port"): it executes the same aten ops the reference's modules execute (nn.Conv2d -> oneDNN,
BatchNorm in eval mode, SiLU, max_pool2d, nearest upsample, cat), layer by layer, on a module tree built by
drone_yolo_b200.nn.tasks (same structure and state_dict as the reference's).

Cited reference forwards:
  Conv.forward / forward_fuse      ultralytics/nn/modules/conv.py:49-55
  DWConv                           conv.py:102-107
  RepConv.forward                  conv.py:202-205
  RepVGGBlock.forward              nn/modules/block.py:1480-1490 (un-merged, as BaseModel.fuse leaves it — SURVEY F5)
  Bottleneck.forward               block.py:348-350
  C2f.forward                      block.py:238-242
  SPPF.forward                     block.py:187-191
  Concat.forward                   conv.py:331-333
  Detect.forward                   nn/modules/head.py:64-74  (+ decode: oracle/decode_np.py)
  BaseModel._predict_once          nn/tasks.py:134-161
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import decode_np


def _bn(x, bn):
    return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)


def _conv2d(x, c):
    return F.conv2d(x, c.weight, c.bias, c.stride, c.padding, c.dilation, c.groups)


def _act(m, x):
    name = type(m).__name__
    if name == "SiLU":
        return F.silu(x)
    if name == "Identity":
        return x
    raise NotImplementedError(name)


def conv_forward(m, x):
    y = _conv2d(x, m.conv)
    if hasattr(m, "bn"):
        y = _bn(y, m.bn)
    return _act(m.act, y)


def repvgg_forward(m, x):
    if hasattr(m, "rbr_reparam"):
        return F.silu(_conv2d(x, m.rbr_reparam))
    idt = 0 if m.rbr_identity is None else _bn(x, m.rbr_identity)
    dense = _bn(_conv2d(x, m.rbr_dense.conv), m.rbr_dense.bn)
    one = _bn(_conv2d(x, m.rbr_1x1.conv), m.rbr_1x1.bn)
    return F.silu(dense + one + idt)


def repconv_forward(m, x):
    if hasattr(m, "conv"):
        return _act(m.act, _conv2d(x, m.conv))
    idt = 0 if m.bn is None else _bn(x, m.bn)
    return _act(m.act, conv_forward(m.conv1, x) + conv_forward(m.conv2, x) + idt)


def bottleneck_forward(m, x):
    y = conv_forward(m.cv2, conv_forward(m.cv1, x))
    return x + y if m.add else y


def c2f_forward(m, x):
    y = list(conv_forward(m.cv1, x).chunk(2, 1))
    for b in m.m:
        y.append(bottleneck_forward(b, y[-1]))
    return conv_forward(m.cv2, torch.cat(y, 1))


def sppf_forward(m, x):
    y = [conv_forward(m.cv1, x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], m.k, 1, m.k // 2))
    return conv_forward(m.cv2, torch.cat(y, 1))


def detect_raw(m, xs):
    outs = []
    for i in range(m.nl):
        a = xs[i]
        for layer in m.cv2[i]:
            a = conv_forward(layer, a) if hasattr(layer, "conv") else _conv2d(a, layer)
        b = xs[i]
        for layer in m.cv3[i]:
            b = conv_forward(layer, b) if hasattr(layer, "conv") else _conv2d(b, layer)
        outs.append(torch.cat((a, b), 1))
    return outs


def module_forward(m, x):
    name = type(m).__name__
    if name == "Sequential":
        for sub in m:
            x = module_forward(sub, x)
        return x
    if name in ("Conv", "DWConv"):
        return conv_forward(m, x)
    if name == "RepVGGBlock":
        return repvgg_forward(m, x)
    if name == "RepConv":
        return repconv_forward(m, x)
    if name == "C2f":
        return c2f_forward(m, x)
    if name == "SPPF":
        return sppf_forward(m, x)
    if name == "Upsample":
        return F.interpolate(x, scale_factor=2.0, mode="nearest")
    if name == "Concat":
        return torch.cat(x, 1)
    raise NotImplementedError(name)


@torch.no_grad()
def forward(model, x, return_features=False):
    """model: drone_yolo_b200 DetectionModel ON CPU (fp32); x: (B,3,H,W) fp32 in [0,1].
    Returns (y (B,4+nc,A) float32 ndarray, [raw maps (B,no,H_l,W_l) tensors])."""
    y = []
    feats = {}
    for m in model.model:
        if m.f != -1:
            x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
        if type(m).__name__ == "Detect":
            raw = detect_raw(m, x)
            dec = decode_np.decode([r.numpy() for r in raw], [float(s) for s in m.stride.tolist()], m.nc)
            return (dec, raw, feats) if return_features else (dec, raw)
        x = module_forward(m, x)
        y.append(x if m.i in model.save else None)
        if return_features:
            feats[m.i] = x
    raise RuntimeError("model has no Detect head")


@torch.no_grad()
def forward_raw(model, x):
    """The conv stack only: the per-level raw head maps (B, no, H_l, W_l) (Detect.forward before _inference, head.py:64-71)."""
    y = []
    for m in model.model:
        if m.f != -1:
            x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
        if type(m).__name__ == "Detect":
            return detect_raw(m, x)
        x = module_forward(m, x)
        y.append(x if m.i in model.save else None)
    raise RuntimeError("model has no Detect head")


def fuse_like_reference(model):
    """What the reference's BaseModel.fuse does on this graph (nn/tasks.py:193-221): Conv+BN folded, RepVGGBlock left
    as two conv+BN branches (it never calls switch_to_deploy).  Used by the CPU baseline timing."""
    for m in model.modules():
        if type(m).__name__ in ("Conv", "DWConv") and hasattr(m, "bn"):
            m.fuse()
    return model
