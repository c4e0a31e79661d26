"""Convolution modules of the Drone-YOLO graphs, same constructor signatures and state_dict keys as the
reference (ultralytics/nn/modules/conv.py), executed by the sm_100a kernels of libdroneyolo.

`forward` accepts the reference's (B, C, H, W) tensors (any dtype/layout on a CUDA device), and returns a bf16
channels-last tensor.  There is no PyTorch compute path: BN is folded in fp32 on the host
(`fused_weight_bias`), the result is packed to bf16 once and the conv runs as a tcgen05 implicit GEMM.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ... import kernels as K

__all__ = ("autopad", "fold_bn", "Conv", "DWConv", "RepConv", "Concat")


def autopad(k, p=None, d=1):
    """'same' padding for kernel k and dilation d (reference conv.py:28-34)."""
    if d > 1:
        k = d * (k - 1) + 1 if isinstance(k, int) else [d * (x - 1) + 1 for x in k]
    if p is None:
        p = k // 2 if isinstance(k, int) else [x // 2 for x in k]
    return p


def fold_bn(weight: torch.Tensor, bias, bn: nn.BatchNorm2d):
    """W' = diag(g/sqrt(var+eps)) W, b' = beta + (b - mean) g/sqrt(var+eps), in fp32
    (reference utils/torch_utils.py:242-269 and conv.py:221-247)."""
    w = weight.detach().float()
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    b0 = torch.zeros(w.shape[0], device=w.device) if bias is None else bias.detach().float()
    return w * scale.reshape(-1, 1, 1, 1), bn.bias.detach().float() + (b0 - bn.running_mean.detach().float()) * scale


def as_input(x: torch.Tensor) -> torch.Tensor:
    """Bring a reference-style activation to what the kernels read: bf16, channels-last, C % 8 == 0."""
    if x.dtype == torch.bfloat16 and x.dim() == 4 and x.shape[1] % 8 == 0:
        try:
            K.nhwc_view(x)
            return x
        except K._C.DroneYoloError:
            pass
    K._C.require_cuda(x)
    return K.to_nhwc_bf16(x)


class _PackedMixin:
    """Caches the folded + packed weights; `invalidate()` after any weight change."""

    def invalidate(self):
        self.__dict__.pop("_packed", None)

    def _load_from_state_dict(self, *args, **kwargs):  # noqa: D401 - nn.Module hook
        self.invalidate()
        return super()._load_from_state_dict(*args, **kwargs)

    def _apply(self, fn, *a, **k):
        self.invalidate()
        return super()._apply(fn, *a, **k)


class Conv(_PackedMixin, nn.Module):
    """Conv2d(no bias) + BatchNorm2d + SiLU  (reference conv.py:37-55)."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, autopad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2)
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()

    # ---- reparameterisation -------------------------------------------------------------------
    def fused_weight_bias(self):
        if hasattr(self, "bn"):
            return fold_bn(self.conv.weight, self.conv.bias, self.bn)
        b = self.conv.bias
        return self.conv.weight.detach().float(), (torch.zeros(self.conv.out_channels, device=self.conv.weight.device)
                                                   if b is None else b.detach().float())

    def fuse(self):
        """In-place Conv+BN fold, as BaseModel.fuse does for Conv (reference tasks.py:206-209)."""
        if hasattr(self, "bn"):
            w, b = self.fused_weight_bias()
            conv = nn.Conv2d(self.conv.in_channels, self.conv.out_channels, self.conv.kernel_size, self.conv.stride,
                             self.conv.padding, self.conv.dilation, self.conv.groups, bias=True).requires_grad_(False)
            conv = conv.to(w.device)
            conv.weight.copy_(w)
            conv.bias.copy_(b)
            self.conv = conv
            delattr(self, "bn")
            self.forward = self.forward_fuse
            self.invalidate()
        return self

    # ---- execution ------------------------------------------------------------------------------
    def _check_supported(self):
        c = self.conv
        k, s = c.kernel_size[0], c.stride[0]
        if c.groups != 1 or c.dilation != (1, 1) or c.kernel_size[0] != c.kernel_size[1] or k not in (1, 3) or \
                s not in (1, 2) or c.padding != (k // 2, k // 2) or (k == 1 and s != 1):
            raise K._C.DroneYoloError(f"Conv {c} is outside what the sm_100a conv kernel implements")
        if not isinstance(self.act, (nn.SiLU, nn.Identity)):
            raise K._C.DroneYoloError(f"activation {self.act} is not implemented (SiLU / Identity only)")

    @property
    def is_stem(self):
        """model.0: 3-channel image in, 3x3 stride 2 — runs on the dedicated stem kernel (fp32 NCHW input)."""
        c = self.conv
        return c.in_channels == 3 and c.kernel_size == (3, 3) and c.stride == (2, 2) and isinstance(self.act, nn.SiLU)

    def packed(self):
        if "_packed" not in self.__dict__:
            self._check_supported()
            w, b = self.fused_weight_bias()
            if self.is_stem:
                self.__dict__["_packed"] = (w.reshape(w.shape[0], 27).contiguous(), b.contiguous())
            else:
                self.__dict__["_packed"] = K.pack_conv_weight(w, b)
        return self.__dict__["_packed"]

    def run(self, x, out=None, residual=None, out_dtype=torch.bfloat16):
        """act(conv(x)+b) [+ residual] written to `out` (may be a channel slice of a concat buffer)."""
        if self.training:
            raise K._C.DroneYoloError("drone_yolo_b200 implements the inference path only: call .eval()")
        w, b = self.packed()
        c = self.conv
        if self.is_stem:
            K._C.require_cuda(x)
            x = x.contiguous() if x.dtype == torch.uint8 else x.float().contiguous()
            return K.stem_conv(x, w, b, out=out)
        return K.conv2d(as_input(x), w, b, c.out_channels, c.kernel_size[0], c.stride[0], isinstance(self.act, nn.SiLU),
                        residual=residual, out=out, out_dtype=out_dtype)

    def forward(self, x):
        return self.run(x)

    def forward_fuse(self, x):
        return self.run(x)


class DWConv(Conv):
    """Depth-wise style conv, groups = gcd(c1, c2)  (reference conv.py:102-107)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)

    def _check_supported(self):
        c = self.conv
        if not (c.kernel_size == (3, 3) and c.stride == (2, 2) and c.padding == (1, 1) and c.dilation == (1, 1)
                and c.groups == c.out_channels and c.in_channels == 2 * c.out_channels and isinstance(self.act, nn.SiLU)):
            raise K._C.DroneYoloError(f"DWConv {c}: only the -sf graph form (k3 s2, c1 == 2*c2, groups == c2) is implemented")

    def packed(self):
        if "_packed" not in self.__dict__:
            self._check_supported()
            w, b = self.fused_weight_bias()
            self.__dict__["_packed"] = (w.contiguous(), b.contiguous())
        return self.__dict__["_packed"]

    def run(self, x, out=None, residual=None, out_dtype=torch.bfloat16):
        if self.training:
            raise K._C.DroneYoloError("drone_yolo_b200 implements the inference path only: call .eval()")
        w, b = self.packed()
        return K.dwconv3x3s2(as_input(x), w, b, out=out)


class RepConv(_PackedMixin, nn.Module):
    """RepVGG-style block: 3x3 conv+BN, 1x1 conv+BN (+ identity BN), summed, SiLU  (reference conv.py:174-275).

    The kernels always run the re-parameterised single 3x3 conv; `fuse_convs()` additionally rewrites the module
    like the reference does."""

    default_act = nn.SiLU()

    def __init__(self, c1, c2, k=3, s=1, p=1, g=1, d=1, act=True, bn=False, deploy=False):
        super().__init__()
        assert k == 3 and p == 1
        self.g, self.c1, self.c2 = g, c1, c2
        self.act = self.default_act if act is True else act if isinstance(act, nn.Module) else nn.Identity()
        self.bn = nn.BatchNorm2d(num_features=c1) if bn and c2 == c1 and s == 1 else None
        self.conv1 = Conv(c1, c2, k, s, p=p, g=g, act=False)
        self.conv2 = Conv(c1, c2, 1, s, p=(p - k // 2), g=g, act=False)

    def _fuse_bn_tensor(self, branch):
        if branch is None:
            return 0, 0
        if isinstance(branch, Conv):
            return fold_bn(branch.conv.weight, None, branch.bn)
        # identity branch: a BatchNorm applied to x == 3x3 conv with a centred one-hot kernel
        input_dim = self.c1 // self.g
        ident = torch.zeros((self.c1, input_dim, 3, 3), device=branch.weight.device)
        for i in range(self.c1):
            ident[i, i % input_dim, 1, 1] = 1.0
        return fold_bn(ident, None, branch)

    def get_equivalent_kernel_bias(self):
        """K = K3 + pad(K1) + K_id, b = b3 + b1 + b_id  (reference conv.py:206-211)."""
        if hasattr(self, "conv"):
            return self.conv.weight.detach().float(), self.conv.bias.detach().float()
        k3, b3 = self._fuse_bn_tensor(self.conv1)
        k1, b1 = self._fuse_bn_tensor(self.conv2)
        kid, bid = self._fuse_bn_tensor(self.bn)
        return k3 + torch.nn.functional.pad(k1, [1, 1, 1, 1]) + kid, b3 + b1 + bid

    fused_weight_bias = get_equivalent_kernel_bias

    def fuse_convs(self):
        """Collapse the branches into one Conv2d named `conv` (reference conv.py:249-275)."""
        if hasattr(self, "conv"):
            return
        kernel, bias = self.get_equivalent_kernel_bias()
        c = self.conv1.conv
        self.conv = nn.Conv2d(c.in_channels, c.out_channels, c.kernel_size, c.stride, c.padding, c.dilation, c.groups,
                              bias=True).requires_grad_(False).to(kernel.device)
        self.conv.weight.data = kernel
        self.conv.bias.data = bias
        for para in self.parameters():
            para.detach_()
        for name in ("conv1", "conv2", "nm", "bn", "id_tensor"):
            if hasattr(self, name):
                self.__delattr__(name)
        self.invalidate()

    def _geom(self):
        c = self.conv if hasattr(self, "conv") else self.conv1.conv
        if c.groups != 1 or c.stride[0] not in (1, 2):
            raise K._C.DroneYoloError("RepConv with groups != 1 is not implemented")
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
        return K.conv2d(as_input(x), w, b, cout, 3, s, isinstance(self.act, nn.SiLU), residual=residual, out=out)

    def forward(self, x):
        return self.run(x)

    forward_fuse = forward


class Concat(nn.Module):
    """Channel concat (reference conv.py:323-333).  Stand-alone it is a device copy into one NHWC buffer; inside a
    compiled plan (engine/plan.py) the producers write their slices directly and this op disappears."""

    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension

    def forward(self, x):
        if self.d != 1:
            raise K._C.DroneYoloError("Concat: only the channel dimension is implemented")
        xs = [as_input(t) for t in x]
        B, _, H, W = xs[0].shape
        out = K.empty_nhwc(B, sum(t.shape[1] for t in xs), H, W, xs[0].device)
        c0 = 0
        for t in xs:
            out[:, c0:c0 + t.shape[1]].copy_(t)
            c0 += t.shape[1]
        return out
