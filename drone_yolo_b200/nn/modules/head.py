"""Detect head (reference ultralytics/nn/modules/head.py:21-172, legacy=True branch used by the v8 YAMLs).

Per level the two branches' first 3x3 convs read the same input and run as ONE implicit GEMM (N = c2 + c3); the
final 1x1 convs write fp32 logits into one NHWC buffer [B,H,W,64+ceil16(nc)] whose first 64+nc channels are the
reference's raw map `x[i]`; decode (DFL + dist2bbox + sigmoid) is one fused kernel over all levels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import kernels as K
from .block import DFL
from .conv import Conv, as_input

__all__ = ("Detect",)


class Detect(nn.Module):
    """YOLO Detect head for detection models."""

    dynamic = False
    export = False
    format = None
    end2end = False
    max_det = 300
    shape = None
    anchors = torch.empty(0)
    strides = torch.empty(0)
    legacy = False  # parse_model sets True for the v8 graphs (reference tasks.py:1061-1062)

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        if not self.legacy:
            raise K._C.DroneYoloError("Detect(legacy=False) (the YOLO11 DWConv class branch) is not part of Drone-YOLO")
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch
        )
        self.cv3 = nn.ModuleList(nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, self.nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max) if self.reg_max > 1 else nn.Identity()

    # ---- packed weights -------------------------------------------------------------------------
    def invalidate(self):
        self.__dict__.pop("_packed", None)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *args, **kwargs):
        self.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    def packed(self):
        """Per level: merged first conv (box|cls), and the two final 1x1 convs."""
        if "_packed" not in self.__dict__:
            levels = []
            for i in range(self.nl):
                wa, ba = self.cv2[i][0].fused_weight_bias()
                wb, bb = self.cv3[i][0].fused_weight_bias()
                first = K.pack_conv_weight(torch.cat((wa, wb), 0), torch.cat((ba, bb), 0))
                box = K.pack_conv_weight(self.cv2[i][2].weight.detach().float(), self.cv2[i][2].bias.detach().float())
                cls = K.pack_conv_weight(self.cv3[i][2].weight.detach().float(), self.cv3[i][2].bias.detach().float())
                levels.append((first, box, cls))
            self.__dict__["_packed"] = levels
        return self.__dict__["_packed"]

    @property
    def raw_ld(self):
        return 4 * self.reg_max + (self.nc + 15) // 16 * 16

    def raw_maps(self, x):
        """Run the conv branches. Returns per level an fp32 NHWC buffer viewed as (B, no, H, W) — the reference's
        `torch.cat((cv2[i](x[i]), cv3[i](x[i])), 1)` (head.py:69-70)."""
        outs = []
        packed = self.packed()
        for i in range(self.nl):
            xi = as_input(x[i])
            B, _, H, W = xi.shape
            c2 = self.cv2[i][0].conv.out_channels
            c3 = self.cv3[i][0].conv.out_channels
            (wf, bf), (wbx, bbx), (wcl, bcl) = packed[i]
            t1 = K.conv2d(xi, wf, bf, c2 + c3, 3, 1, True)                       # [box feat | cls feat]
            t2 = K.empty_nhwc(B, c2 + c3, H, W, xi.device)
            self.cv2[i][1].run(t1[:, :c2], out=t2[:, :c2])
            self.cv3[i][1].run(t1[:, c2:], out=t2[:, c2:])
            raw = K.empty_nhwc(B, self.raw_ld, H, W, xi.device, torch.float32)
            K.conv2d(t2[:, :c2], wbx, bbx, 4 * self.reg_max, 1, 1, False, out=raw[:, :4 * self.reg_max])
            ncp = self.raw_ld - 4 * self.reg_max           # padded class rows are zero weights: pad channels become 0
            K.conv2d(t2[:, c2:], wcl, bcl, ncp, 1, 1, False, out=raw[:, 4 * self.reg_max:])
            outs.append(raw[:, :self.no])
        return outs

    def forward(self, x):
        if self.training:
            raise K._C.DroneYoloError("drone_yolo_b200 implements the inference path only: call .eval()")
        x = self.raw_maps(x)
        y = self._inference(x)
        return y if self.export else (y, x)

    def _inference(self, x):
        """Decode boxes and class probabilities from the raw maps -> (B, 4+nc, A) fp32 (head.py:100-131)."""
        strides = [float(s) for s in self.stride.tolist()]
        if any(s <= 0 for s in strides):
            raise K._C.DroneYoloError("Detect.stride is not set; build the head through DetectionModel")
        return K.detect_decode(list(x), strides, self.nc)

    def bias_init(self):
        """Initialise the output biases (reference head.py:133-144); needs `stride`."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)
        self.invalidate()
