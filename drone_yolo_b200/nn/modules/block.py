"""Block modules of the Drone-YOLO graphs (reference ultralytics/nn/modules/block.py): DFL, SPPF, C2f,
Bottleneck and RepVGGBlock, with the reference's constructor signatures and state_dict keys.

The blocks own their concat buffers: C2f's `torch.cat(y, 1)` and SPPF's `torch.cat` never happen, the convs
write their outputs straight into channel slices of one NHWC buffer (`Conv.run(out=...)`).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import kernels as K
from .conv import Conv, _PackedMixin, as_input, fold_bn

__all__ = ("DFL", "SPPF", "C2f", "Bottleneck", "RepVGGBlock", "conv_bn")


class DFL(nn.Module):
    """Distribution Focal Loss integral (reference block.py:58-77): softmax over c1 bins, expectation with
    weights 0..c1-1.  Kept for state_dict parity (`dfl.conv.weight`); Detect runs it inside the fused decode kernel."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        x = torch.arange(c1, dtype=torch.float)
        self.conv.weight.data[:] = nn.Parameter(x.view(1, c1, 1, 1))
        self.c1 = c1

    def forward(self, x):
        raise K._C.DroneYoloError("DFL runs fused inside dy_detect_decode (Detect._inference); it has no stand-alone kernel")


class SPPF(nn.Module):
    """Spatial Pyramid Pooling - Fast (reference block.py:172-191)."""

    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)
        self.k = k

    def forward(self, x, out=None):
        if self.k != 5:
            raise K._C.DroneYoloError("SPPF: only k=5 is implemented")
        x = as_input(x)
        B, _, H, W = x.shape
        c_ = self.cv1.conv.out_channels
        cat = K.empty_nhwc(B, 4 * c_, H, W, x.device)
        self.cv1.run(x, out=cat[:, :c_])
        K.sppf_pool(cat, c_)
        return self.cv2.run(cat, out=out)


class Bottleneck(nn.Module):
    """Two convs with an optional residual (reference block.py:337-350)."""

    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2

    def forward(self, x, out=None):
        x = as_input(x)
        return self.cv2.run(self.cv1.run(x), out=out, residual=x if self.add else None)


class C2f(nn.Module):
    """CSP bottleneck with 2 convs, "faster" variant (reference block.py:227-249)."""

    def __init__(self, c1, c2, n=1, shortcut=False, g=1, e=0.5):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut, g, k=((3, 3), (3, 3)), e=1.0) for _ in range(n))

    def forward(self, x, out=None):
        x = as_input(x)
        B, _, H, W = x.shape
        c, n = self.c, len(self.m)
        cat = K.empty_nhwc(B, (2 + n) * c, H, W, x.device)       # [a | b | m0(b) | m1(..) | ...]
        self.cv1.run(x, out=cat[:, :2 * c])
        for i, m in enumerate(self.m):
            m(cat[:, (1 + i) * c:(2 + i) * c], out=cat[:, (2 + i) * c:(3 + i) * c])
        return self.cv2.run(cat, out=out)


def conv_bn(in_channels, out_channels, kernel_size, stride, padding, groups=1):
    """Conv2d(no bias) + BatchNorm2d as a Sequential with members `conv`, `bn` (reference block.py:1365-1372)."""
    result = nn.Sequential()
    result.add_module("conv", nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, groups=groups, bias=False))
    result.add_module("bn", nn.BatchNorm2d(num_features=out_channels))
    return result


class RepVGGBlock(_PackedMixin, nn.Module):
    """RepVGG block used by Drone-YOLO for the backbone downsamples (reference block.py:1393-1490):
    SiLU(BN(conv3x3(x)) + BN(conv1x1(x)) [+ BN(x)]).  The kernels always execute the re-parameterised form
    (one 3x3 conv + bias), which `switch_to_deploy()` also materialises like the reference does."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=1, dilation=1, groups=1,
                 padding_mode="zeros", deploy=False, use_se=False):
        super().__init__()
        if use_se:
            raise K._C.DroneYoloError("RepVGGBlock(use_se=True) is not used by the Drone-YOLO graphs and is not implemented")
        if kernel_size != 3 or padding != 1 or dilation != 1 or groups != 1 or padding_mode != "zeros":
            raise K._C.DroneYoloError("RepVGGBlock: only k=3, p=1, d=1, g=1, zero padding is implemented")
        self.deploy = deploy
        self.groups = groups
        self.in_channels = in_channels
        self.nonlinearity = nn.SiLU()
        self.se = nn.Identity()
        if deploy:
            self.rbr_reparam = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias=True)
        else:
            self.rbr_identity = nn.BatchNorm2d(in_channels) if out_channels == in_channels and stride == 1 else None
            self.rbr_dense = conv_bn(in_channels, out_channels, kernel_size, stride, padding, groups)
            self.rbr_1x1 = conv_bn(in_channels, out_channels, 1, stride, padding - kernel_size // 2, groups)

    def _fuse_bn_tensor(self, branch):
        if branch is None:
            return 0, 0
        if isinstance(branch, nn.Sequential):
            return fold_bn(branch.conv.weight, None, branch.bn)
        input_dim = self.in_channels // self.groups
        ident = torch.zeros((self.in_channels, input_dim, 3, 3), device=branch.weight.device)
        for i in range(self.in_channels):
            ident[i, i % input_dim, 1, 1] = 1.0
        return fold_bn(ident, None, branch)

    def get_equivalent_kernel_bias(self):
        """K3 + pad(K1) + K_id and the summed biases (reference block.py:1440-1478)."""
        if hasattr(self, "rbr_reparam"):
            return self.rbr_reparam.weight.detach().float(), self.rbr_reparam.bias.detach().float()
        k3, b3 = self._fuse_bn_tensor(self.rbr_dense)
        k1, b1 = self._fuse_bn_tensor(self.rbr_1x1)
        kid, bid = self._fuse_bn_tensor(self.rbr_identity)
        return k3 + torch.nn.functional.pad(k1, [1, 1, 1, 1]) + kid, b3 + b1 + bid

    fused_weight_bias = get_equivalent_kernel_bias

    def switch_to_deploy(self):
        if hasattr(self, "rbr_1x1"):
            kernel, bias = self.get_equivalent_kernel_bias()
            c = self.rbr_dense.conv
            self.rbr_reparam = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, c.dilation,
                                         c.groups, bias=True).to(kernel.device)
            self.rbr_reparam.weight.data = kernel
            self.rbr_reparam.bias.data = bias
            for para in self.parameters():
                para.detach_()
            self.rbr_dense = self.rbr_reparam
            self.__delattr__("rbr_1x1")
            if hasattr(self, "rbr_identity"):
                self.__delattr__("rbr_identity")
            self.deploy = True
            self.invalidate()

    def _geom(self):
        c = self.rbr_reparam if hasattr(self, "rbr_reparam") else self.rbr_dense.conv
        return c.out_channels, c.stride[0]

    def packed(self):
        if "_packed" not in self.__dict__:
            w, b = self.get_equivalent_kernel_bias()
            self.__dict__["_packed"] = K.pack_conv_weight(w, b)
        return self.__dict__["_packed"]

    def run(self, x, out=None, residual=None, out_dtype=torch.bfloat16):
        if self.training:
            raise K._C.DroneYoloError("drone_yolo_b200 implements the inference path only: call .eval()")
        cout, s = self._geom()
        w, b = self.packed()
        return K.conv2d(as_input(x), w, b, cout, 3, s, True, residual=residual, out=out)

    def forward(self, inputs):
        return self.run(inputs)
