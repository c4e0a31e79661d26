"""Tensor-level wrappers over the C-ABI (drone_yolo_b200/_C.py -> libdroneyolo.so).

Activations are torch tensors of logical shape (B, C, H, W) in channels_last memory (i.e. NHWC bf16), possibly a
channel slice `buf[:, c0:c1]` of a wider concat buffer.  PyTorch is used only for memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _C

BLOCK_K = 64


def _on_tensor_device(fn):
    """The library launches on the CALLING THREAD's current device, with a stream and pointers that belong to the tensors'
    device: make that device current for the duration of the call (predict(device='cuda:1') from a process whose current
    device is 0 otherwise fails with invalid-resource-handle errors)."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor):
                a = a[0]
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        if dev is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)

    return wrapped


def _ceil(a: int, b: int) -> int:
    return (a + b - 1) // b * b


def nhwc_view(t: torch.Tensor, what: str = "tensor"):
    """Validate a (B,C,H,W) channels-last (slice) tensor; return (ptr, ld, B, H, W, C)."""
    _C.require_cuda(t)
    if t.dim() != 4:
        raise _C.DroneYoloError(f"{what}: expected a 4-D (B,C,H,W) tensor, got shape {tuple(t.shape)}")
    B, Cc, H, W = t.shape
    sb, sc, sh, sw = t.stride()
    ld = sw if W > 1 else (sh if H > 1 else max(Cc, sc))
    ok = (sc == 1 or Cc == 1) and (W == 1 or sw == ld) and (H == 1 or sh == W * ld) and (B == 1 or sb == H * W * ld)
    if not ok or ld < Cc:
        raise _C.DroneYoloError(
            f"{what}: expected channels_last (NHWC) memory or a channel slice of it, got shape {tuple(t.shape)} "
            f"strides {t.stride()}"
        )
    return t.data_ptr(), ld, B, H, W, Cc


def empty_nhwc(B: int, Cc: int, H: int, W: int, device, dtype=torch.bfloat16) -> torch.Tensor:
    return torch.empty((B, H, W, Cc), device=device, dtype=dtype).permute(0, 3, 1, 2)


def to_nhwc_bf16(x: torch.Tensor) -> torch.Tensor:
    """(B,C,H,W) any layout/dtype -> bf16 channels_last with C padded to a multiple of 8 kept as a slice."""
    B, Cc, H, W = x.shape
    Cp = _ceil(Cc, 8)
    buf = torch.zeros((B, H, W, Cp), device=x.device, dtype=torch.bfloat16)
    buf[..., :Cc] = x.permute(0, 2, 3, 1)
    return buf.permute(0, 3, 1, 2)[:, :Cc]


def pack_conv_weight(w: torch.Tensor, b: Optional[torch.Tensor]):
    """fp32 [Cout,Cin,k,k] (+bias [Cout]) -> bf16 [k*k, Cout_pad, Cin_pad] tap-major K-contiguous, fp32 bias [Cout_pad].

    BN folding / RepVGG merging must already have been done in fp32 (see nn.modules.*.fused_weight_bias)."""
    Cout, Cin, kh, kw = w.shape
    cout_p, cin_p = _ceil(Cout, 16), _ceil(Cin, BLOCK_K)
    packed = torch.zeros((kh * kw, cout_p, cin_p), device=w.device, dtype=torch.float32)
    packed[:, :Cout, :Cin] = w.float().permute(2, 3, 0, 1).reshape(kh * kw, Cout, Cin)
    bias = torch.zeros((cout_p,), device=w.device, dtype=torch.float32)
    if b is not None:
        bias[:Cout] = b.float()
    return packed.to(torch.bfloat16).contiguous(), bias.contiguous()


def conv_desc(x: torch.Tensor, w_packed: torch.Tensor, bias: torch.Tensor, cout: int, k: int, s: int, act: bool,
              out: Optional[torch.Tensor], residual: Optional[torch.Tensor] = None, up_out: Optional[torch.Tensor] = None,
              tail=None, pre_add: Optional[torch.Tensor] = None) -> _C.ConvDesc:
    """`tail` = (w2_packed, bias2, cout2, out2): fused 1x1 conv behind this one; `out` may then be None.  out2 fp32 = raw
    output conv (no activation; Detect), out2 bf16 = Conv + SiLU (C2f.cv1); a fifth element selects the fused decode."""
    xp, xld, B, H, W, Cin = nhwc_view(x, "conv input")
    if out is None:
        if tail is None:
            raise _C.DroneYoloError("conv needs an output tensor")
    if out is None:
        p_ = k // 2
        op, old, Bo, Ho, Wo, Co = 0, cout, B, (H + 2 * p_ - k) // s + 1, (W + 2 * p_ - k) // s + 1, cout
        out_dtype_is_f32 = False
    else:
        op, old, Bo, Ho, Wo, Co = nhwc_view(out, "conv output")
        out_dtype_is_f32 = out.dtype == torch.float32
    if x.dtype != torch.bfloat16:
        raise _C.DroneYoloError(f"conv input must be bf16, got {x.dtype}")
    p = k // 2
    eh, ew = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    if (Bo, Ho, Wo, Co) != (B, eh, ew, cout):
        raise _C.DroneYoloError(f"conv output shape {(Bo, Co, Ho, Wo)} != expected {(B, cout, eh, ew)}")
    if w_packed.shape != (k * k, _ceil(cout, 16), _ceil(Cin, BLOCK_K)) or w_packed.dtype != torch.bfloat16:
        raise _C.DroneYoloError(f"packed weight shape {tuple(w_packed.shape)} does not match Cin={Cin} Cout={cout} k={k}")
    if out is not None and out.dtype not in (torch.bfloat16, torch.float32):
        raise _C.DroneYoloError(f"conv output dtype {out.dtype} unsupported")
    d = _C.ConvDesc()
    d.in_, d.in_ld, d.B, d.H, d.W, d.Cin = xp, xld, B, H, W, Cin
    d.weight, d.bias = w_packed.data_ptr(), bias.data_ptr()
    d.Cout, d.ksize, d.stride = cout, k, s
    d.out, d.out_ld = op, old
    d.out_dtype = _C.DY_F32 if out_dtype_is_f32 else _C.DY_BF16
    if residual is not None:
        rp, rld, Br, Hr, Wr, Cr = nhwc_view(residual, "conv residual")
        if (Br, Hr, Wr, Cr) != (B, eh, ew, cout) or residual.dtype != torch.bfloat16:
            raise _C.DroneYoloError("conv residual must be bf16 with the output's shape")
        d.residual, d.res_ld = rp, rld
    else:
        d.residual, d.res_ld = None, 0
    d.act = _C.DY_ACT_SILU if act else _C.DY_ACT_NONE
    if up_out is not None:
        up, uld, Bu, Hu, Wu, Cu = nhwc_view(up_out, "conv upsampled output")
        if (Bu, Hu, Wu, Cu) != (B, 2 * eh, 2 * ew, cout) or up_out.dtype != torch.bfloat16 or out is None or out.dtype != torch.bfloat16:
            raise _C.DroneYoloError("conv up_out must be bf16 (B, Cout, 2*Ho, 2*Wo), with a bf16 primary output")
        d.up_out, d.up_ld = up, uld
    else:
        d.up_out, d.up_ld = None, 0
    d.tail_decode, d.y, d.y_A, d.y_nc, d.y_anchor_off, d.y_stride = 0, None, 0, 0, 0, 0.0
    if pre_add is not None:
        # half-resolution fp32 addend in front of the activation: out = act(conv(x) + bias + up2x(pre_add))
        pp, pld, Bp, Hp, Wp, Cp = nhwc_view(pre_add, "conv pre_add")
        if (Bp, 2 * Hp, 2 * Wp, Cp) != (B, eh, ew, cout) or pre_add.dtype != torch.float32:
            raise _C.DroneYoloError("conv pre_add must be fp32 (B, Cout, Ho/2, Wo/2)")
        d.pre_add, d.pre_ld = pp, pld
    else:
        d.pre_add, d.pre_ld = None, 0
    if tail is not None:
        w2, b2, cout2, out2 = tail[:4]
        dec = tail[4] if len(tail) > 4 else None
        if w2.shape != (1, _ceil(cout2, 16), _ceil(cout, BLOCK_K)) or w2.dtype != torch.bfloat16:
            raise _C.DroneYoloError(f"packed tail weight shape {tuple(w2.shape)} does not match {cout}->{cout2}")
        d.weight2, d.bias2, d.Cout2 = w2.data_ptr(), b2.data_ptr(), cout2
        if dec is not None:
            # fused Detect decode: dec = (mode 1 box | 2 class, y (B, 4+nc, A) fp32 contiguous, first anchor of the level, stride)
            mode, y, anchor_off, stride = dec
            _C.require_cuda(y)
            if y.dtype != torch.float32 or not y.is_contiguous() or y.dim() != 3 or y.shape[0] != B:
                raise _C.DroneYoloError("conv tail decode: y must be a contiguous fp32 (B, 4+nc, A) tensor")
            d.out2, d.out2_ld = None, 0
            d.tail_decode, d.y, d.y_A, d.y_nc, d.y_anchor_off, d.y_stride = mode, y.data_ptr(), y.shape[2], y.shape[1] - 4, anchor_off, float(stride)
        else:
            o2, o2ld, B2, H2, W2, C2 = nhwc_view(out2, "conv tail output")
            if (B2, H2, W2, C2) != (B, eh, ew, cout2) or out2.dtype not in (torch.float32, torch.bfloat16):
                raise _C.DroneYoloError("conv tail output must be fp32 (raw logits) or bf16 (SiLU tail) (B, Cout2, Ho, Wo)")
            d.out2, d.out2_ld = o2, o2ld
            if out2.dtype == torch.bfloat16:        # hidden Conv + SiLU as the tail (dy_conv_desc.tail_decode == 3)
                d.tail_decode = 3
    else:
        d.weight2, d.bias2, d.Cout2, d.out2, d.out2_ld = None, None, 0, None, 0
    return d


@_on_tensor_device
def conv2d(x, w_packed, bias, cout: int, k: int, s: int, act: bool = True, residual=None, out=None,
           out_dtype=torch.bfloat16, up_out=None, tail=None, pre_add=None) -> torch.Tensor:
    """act(conv(x) + bias) [+ residual] on tcgen05 tensor cores (dy_conv2d).  With `tail` (see conv_desc) the fused
    1x1 output conv's fp32 tensor is returned instead and the intermediate is never written."""
    B, _, H, W = x.shape
    p = k // 2
    if out is None and tail is None:
        out = empty_nhwc(B, cout, (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1, x.device, out_dtype)
    d = conv_desc(x, w_packed, bias, cout, k, s, act, out, residual, up_out, tail, pre_add)
    _C.check(_C.lib().dy_conv2d(C.byref(d), _C.stream_ptr(x.device)), "dy_conv2d")
    if tail is not None:
        return tail[4][1] if len(tail) > 4 and tail[4] is not None else tail[3]
    return out


@_on_tensor_device
def stem_conv(x: torch.Tensor, w27: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x: NCHW contiguous [B,3,H,W], fp32 in [0,1] or uint8 0..255 (scaled by 1/255 in the kernel); w27 fp32 [Cout,27];
    returns bf16 NHWC [B,Cout,H/2,W/2] view."""
    _C.require_cuda(x, w27, bias)
    if x.dtype not in (torch.float32, torch.uint8) or not x.is_contiguous() or x.shape[1] != 3:
        raise _C.DroneYoloError("stem_conv expects a contiguous fp32 or uint8 NCHW tensor with 3 channels")
    B, _, H, W = x.shape
    cout = w27.shape[0]
    if out is None:
        out = empty_nhwc(B, cout, (H - 1) // 2 + 1, (W - 1) // 2 + 1, x.device)
    op, old, *_ = nhwc_view(out, "stem output")
    in_dt = _C.DY_U8 if x.dtype == torch.uint8 else _C.DY_F32
    _C.check(_C.lib().dy_stem_conv(x.data_ptr(), in_dt, B, H, W, w27.data_ptr(), bias.data_ptr(), cout, op, old,
                                   _C.stream_ptr(x.device)), "dy_stem_conv")
    return out


@_on_tensor_device
def letterbox_u8(src: torch.Tensor, dst: torch.Tensor, new_w: int, new_h: int, left: int, top: int, fill: int = 114) -> torch.Tensor:
    """One raw frame (h, w, 3) uint8 BGR on the device -> one (3, H, W) uint8 RGB image of the engine's input batch:
    LetterBox (data/augment.py:1544-1610) + BGR->RGB + HWC->CHW in one kernel."""
    _C.require_cuda(src)
    _C.require_cuda(dst)
    if src.dtype != torch.uint8 or dst.dtype != torch.uint8 or src.dim() != 3 or src.shape[2] != 3 or src.stride(2) != 1 or src.stride(1) != 3:
        raise _C.DroneYoloError("letterbox: src must be a uint8 (h, w, 3) image with packed pixels")
    if dst.dim() != 3 or dst.shape[0] != 3 or not dst.is_contiguous():
        raise _C.DroneYoloError("letterbox: dst must be a contiguous uint8 (3, H, W) tensor")
    _C.check(_C.lib().dy_letterbox_u8(src.data_ptr(), src.shape[0], src.shape[1], src.stride(0), dst.data_ptr(), dst.shape[1],
                                      dst.shape[2], int(new_w), int(new_h), int(left), int(top), int(fill), _C.stream_ptr(dst.device)),
             "dy_letterbox_u8")
    return dst


@_on_tensor_device
def letterbox_u8_batch(src: torch.Tensor, dst: torch.Tensor, new_w: int, new_h: int, left: int, top: int, fill: int = 114) -> torch.Tensor:
    """n raw frames of one shape (n, h, w, 3) uint8 BGR on the device -> (n, 3, H, W) uint8 RGB, one launch (see letterbox_u8)."""
    _C.require_cuda(src)
    _C.require_cuda(dst)
    if src.dtype != torch.uint8 or dst.dtype != torch.uint8 or src.dim() != 4 or src.shape[3] != 3 or src.stride(3) != 1 or src.stride(2) != 3:
        raise _C.DroneYoloError("letterbox: src must be a uint8 (n, h, w, 3) batch with packed pixels")
    if dst.dim() != 4 or dst.shape[1] != 3 or dst.shape[0] != src.shape[0] or not dst[0].is_contiguous():
        raise _C.DroneYoloError("letterbox: dst must be a uint8 (n, 3, H, W) batch of contiguous images")
    _C.check(_C.lib().dy_letterbox_u8_batch(src.data_ptr(), src.shape[0], src.stride(0), src.shape[1], src.shape[2], src.stride(1),
                                            dst.data_ptr(), dst.stride(0), dst.shape[2], dst.shape[3], int(new_w), int(new_h),
                                            int(left), int(top), int(fill), _C.stream_ptr(dst.device)), "dy_letterbox_u8_batch")
    return dst


@_on_tensor_device
def sppf_pool(buf: torch.Tensor, c: int) -> torch.Tensor:
    """buf: bf16 NHWC (B,>=4c,H,W); fills channels [c,4c) with mp5, mp5∘mp5, mp5∘mp5∘mp5 of channels [0,c)."""
    p, ld, B, H, W, Cc = nhwc_view(buf, "sppf buffer")
    if Cc < 4 * c or buf.dtype != torch.bfloat16:
        raise _C.DroneYoloError("sppf_pool needs a bf16 buffer with at least 4*c channels")
    _C.check(_C.lib().dy_sppf_pool(p, B, H, W, c, ld, _C.stream_ptr(buf.device)), "dy_sppf_pool")
    return buf


@_on_tensor_device
def upsample2x(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    xp, xld, B, H, W, Cc = nhwc_view(x, "upsample input")
    if out is None:
        out = empty_nhwc(B, Cc, 2 * H, 2 * W, x.device)
    op, old, Bo, Ho, Wo, Co = nhwc_view(out, "upsample output")
    if (Bo, Ho, Wo, Co) != (B, 2 * H, 2 * W, Cc) or x.dtype != torch.bfloat16 or out.dtype != torch.bfloat16:
        raise _C.DroneYoloError("upsample2x: output must be bf16 (B,C,2H,2W)")
    _C.check(_C.lib().dy_upsample2x(xp, xld, B, H, W, Cc, op, old, _C.stream_ptr(x.device)), "dy_upsample2x")
    return out


@_on_tensor_device
def dwconv3x3s2(x: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x bf16 NHWC (B,2c,H,W); w fp32 [c,2,3,3] (BN folded); returns SiLU(conv+bias) (B,c,H/2,W/2)."""
    xp, xld, B, H, W, Cin = nhwc_view(x, "dwconv input")
    cout = w.shape[0]
    if out is None:
        out = empty_nhwc(B, cout, (H - 1) // 2 + 1, (W - 1) // 2 + 1, x.device)
    op, old, *_ = nhwc_view(out, "dwconv output")
    w = w.reshape(cout, 18).float().contiguous()
    _C.check(_C.lib().dy_dwconv3x3s2(xp, xld, B, H, W, Cin, w.data_ptr(), bias.data_ptr(), cout, op, old,
                                     _C.stream_ptr(x.device)), "dy_dwconv3x3s2")
    return out


def decode_desc(levels: Sequence[torch.Tensor], strides: Sequence[float], nc: int, out: torch.Tensor,
                anchor_off: Optional[Sequence[int]] = None) -> _C.DecodeDesc:
    """levels: per-level raw maps (B, 64+nc, H, W); contiguous NCHW or channels-last (slice), fp32 or bf16."""
    d = _C.DecodeDesc()
    no = 64 + nc
    nl = len(levels)
    if not 1 <= nl <= 4:
        raise _C.DroneYoloError("decode supports 1..4 levels")
    layouts, dts = set(), set()
    for i, t in enumerate(levels):
        _C.require_cuda(t)
        B, Cc, H, W = t.shape
        if Cc != no:
            raise _C.DroneYoloError(f"decode level {i}: expected {no} channels, got {Cc}")
        dts.add(t.dtype)
        if t.is_contiguous() and not (H * W > 1 and t.stride(1) == 1):
            layouts.add(_C.DY_NCHW)
            d.ld[i] = 0
        else:
            _, ld, *_ = nhwc_view(t, f"decode level {i}")
            layouts.add(_C.DY_NHWC)
            d.ld[i] = ld
        d.lvl[i], d.H[i], d.W[i], d.stride[i] = t.data_ptr(), H, W, float(strides[i])
    if len(layouts) != 1 or len(dts) != 1:
        raise _C.DroneYoloError("decode: all levels must share one layout and dtype")
    dt = dts.pop()
    if dt not in (torch.bfloat16, torch.float32):
        raise _C.DroneYoloError(f"decode: unsupported dtype {dt}")
    d.nl, d.B, d.nc = nl, levels[0].shape[0], nc
    d.dtype = _C.DY_F32 if dt == torch.float32 else _C.DY_BF16
    d.layout = layouts.pop()
    d.out = out.data_ptr()
    d.A_total = 0
    if anchor_off is not None:                      # some levels are decoded elsewhere (fused Detect tails): explicit placement
        d.A_total = out.shape[2]
        for i, a in enumerate(anchor_off):
            d.anchor_off[i] = int(a)
    return d


@_on_tensor_device
def detect_decode(levels: Sequence[torch.Tensor], strides: Sequence[float], nc: int,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Detect._inference on the GPU: returns y (B, 4+nc, A) fp32."""
    B = levels[0].shape[0]
    A = sum(int(t.shape[2]) * int(t.shape[3]) for t in levels)
    if out is None:
        out = torch.empty((B, 4 + nc, A), device=levels[0].device, dtype=torch.float32)
    elif out.shape != (B, 4 + nc, A) or out.dtype != torch.float32 or not out.is_contiguous():
        raise _C.DroneYoloError("decode: out must be contiguous fp32 (B,4+nc,A)")
    d = decode_desc(levels, strides, nc, out)
    _C.check(_C.lib().dy_detect_decode(C.byref(d), _C.stream_ptr(out.device)), "dy_detect_decode")
    return out


def normalize_classes(classes):
    """The reference's `classes` argument (cfg default None; ops.py:237-238 turns a list, an int or a tensor into a tensor):
    None -> None (no filter), anything else -> list of ints (an empty list filters every detection out)."""
    if classes is None:
        return None
    if isinstance(classes, torch.Tensor):
        classes = classes.detach().cpu().reshape(-1).tolist()
    elif hasattr(classes, "tolist") and not isinstance(classes, (list, tuple)):      # numpy arrays and scalars
        classes = classes.tolist()
    if isinstance(classes, (int, float)):
        classes = [classes]
    return [int(c) for c in classes]


class NmsBuffers:
    """Output + workspace buffers of one dy_nms call shape (reused across calls)."""

    def __init__(self, B: int, nc: int, A: int, max_det: int, multi_label: bool, device):
        self.key = (B, nc, A, max_det, bool(multi_label), str(device))
        nbytes = _C.lib().dy_nms_workspace_bytes(B, nc, A, int(bool(multi_label)))
        self.workspace = torch.empty((nbytes,), device=device, dtype=torch.uint8)
        self.out = torch.zeros((B, max_det, 6), device=device, dtype=torch.float32)
        self.counts = torch.zeros((B,), device=device, dtype=torch.int32)
        self.kept = torch.zeros((B, max_det), device=device, dtype=torch.int64)
        self._classes = None


def nms_desc(pred: torch.Tensor, bufs: NmsBuffers, conf_thres: float, iou_thres: float, max_det: int, max_nms: int,
             max_wh: float, agnostic: bool, multi_label: bool, classes, in_place: bool, want_kept: bool = True,
             rescale: Optional[torch.Tensor] = None) -> _C.NmsDesc:
    _C.require_cuda(pred)
    if pred.dim() != 3 or pred.dtype != torch.float32 or not pred.is_contiguous():
        raise _C.DroneYoloError("nms: prediction must be a contiguous fp32 (B, 4+nc, A) tensor")
    B, ch, A = pred.shape
    nc = ch - 4
    d = _C.NmsDesc()
    d.pred, d.B, d.nc, d.A = pred.data_ptr(), B, nc, A
    d.conf_thres, d.iou_thres = float(conf_thres), float(iou_thres)
    d.max_det, d.max_nms, d.max_wh = int(max_det), int(max_nms), float(max_wh)
    d.agnostic, d.multi_label = int(bool(agnostic)), int(bool(multi_label))
    classes = normalize_classes(classes)
    if classes is not None:
        if len(classes) == 0:
            classes = [nc]                  # no class passes (ops.py:294-295 with an empty tensor): an id outside [0, nc) sets no bit
        arr = (C.c_int32 * len(classes))(*[int(c) for c in classes])
        bufs._classes = arr  # keep alive
        d.classes_host, d.n_classes = C.cast(arr, C.POINTER(C.c_int32)), len(classes)
    else:
        d.classes_host, d.n_classes = None, 0
    d.xyxy_in_place = int(bool(in_place))
    d.out, d.counts = bufs.out.data_ptr(), bufs.counts.data_ptr()
    d.kept = bufs.kept.data_ptr() if want_kept else None
    d.workspace, d.workspace_bytes = bufs.workspace.data_ptr(), bufs.workspace.numel()
    if rescale is not None:
        if rescale.dtype != torch.float32 or tuple(rescale.shape) != (B, 8) or not rescale.is_contiguous() or rescale.device != pred.device:
            raise _C.DroneYoloError("nms: rescale must be a contiguous fp32 (B, 8) tensor on the prediction's device")
        d.rescale = rescale.data_ptr()
    else:
        d.rescale = None
    return d


@_on_tensor_device
def nms(pred: torch.Tensor, conf_thres: float, iou_thres: float, max_det: int = 300, max_nms: int = 30000,
        max_wh: float = 7680, agnostic: bool = False, multi_label: bool = False, classes=None, in_place: bool = False,
        bufs: Optional[NmsBuffers] = None, rescale: Optional[torch.Tensor] = None):
    """Batched NMS on the GPU. Returns (out (B,max_det,6), counts (B,), kept (B,max_det)) device tensors."""
    B, ch, A = pred.shape
    nc = ch - 4
    ml = bool(multi_label) and nc > 1
    if bufs is None or bufs.key != (B, nc, A, max_det, ml, str(pred.device)):
        bufs = NmsBuffers(B, nc, A, max_det, ml, pred.device)
    d = nms_desc(pred, bufs, conf_thres, iou_thres, max_det, max_nms, max_wh, agnostic, ml, classes, in_place, rescale=rescale)
    _C.check(_C.lib().dy_nms(C.byref(d), _C.stream_ptr(pred.device)), "dy_nms")
    return bufs.out, bufs.counts, bufs.kept


@_on_tensor_device
def box_nms_f64(rows: torch.Tensor, iou_thres: float, class_agnostic: bool = False) -> torch.Tensor:
    """Merge NMS of the tiled-frame dispatch (the overlap filter of supervision.InferenceSlicer, mix6.py:84-89):
    rows (n, 6) float64 [x1, y1, x2, y2, conf, cls] on the device -> keep mask (n,) bool in the original row order."""
    _C.require_cuda(rows)
    if rows.dtype != torch.float64 or rows.dim() != 2 or rows.shape[1] != 6 or not rows.is_contiguous():
        raise _C.DroneYoloError("box_nms_f64: rows must be a contiguous float64 (n, 6) tensor")
    n = rows.shape[0]
    keep = torch.zeros((n,), dtype=torch.uint8, device=rows.device)
    if n == 0:
        return keep.bool()
    ws = torch.empty((_C.lib().dy_box_nms_f64_workspace_bytes(n) + 7) // 8, dtype=torch.int64, device=rows.device)
    _C.check(_C.lib().dy_box_nms_f64(rows.data_ptr(), n, float(iou_thres), int(bool(class_agnostic)), keep.data_ptr(),
                                     ws.data_ptr(), ws.numel() * 8, _C.stream_ptr(rows.device)), "dy_box_nms_f64")
    return keep.bool()


def selftest_umma(n: int, k: int) -> float:
    err = C.c_float(float("nan"))
    _C.check(_C.lib().dy_selftest_umma(n, k, C.byref(err), _C.stream_ptr()), "dy_selftest_umma")
    return float(err.value)
