"""Layer plan: compile a DetectionModel into a recorded dy_program for a fixed (micro-batch, H, W).

Replaces the reference's Python layer loop `BaseModel._predict_once` (ultralytics/nn/tasks.py:134-161) and the
chunk/cat/upsample copies around it:
  * every Concat (conv.py:331-333), C2f's chunk+cat (block.py:240-242) and SPPF's cat (block.py:191) disappear:
    producers write straight into channel slices of one NHWC buffer;
  * Detect's two first 3x3 convs per level run as one GEMM (N = c2 + c3);
  * activations live in one arena with lifetime-based reuse, sized for a MICRO-batch so that a layer's output is
    still in the 126 MB L2 when the next layer reads it; the program is replayed per micro-batch over the resident
    batch (stem input / decode output pointers are offset per replay);
  * the whole step (all replays + NMS) is captured in one CUDA graph.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import namedtuple
from typing import Optional

import torch
import torch.nn as nn

from .. import _C
from .. import kernels as K
from ..nn.modules import C2f, Concat, Conv, Detect, DWConv, RepConv, RepVGGBlock, SPPF
from ..nn.tasks import Upsample

Ref = namedtuple("Ref", "buf c0 c H W")     # channel slice [c0, c0+c) of buffer `buf` at resolution HxW


class _Buf:
    __slots__ = ("C", "H", "W", "esz", "first", "last", "offset", "nbytes", "pinned")

    def __init__(self, Cc, H, W, esz, first):
        self.C, self.H, self.W, self.esz, self.first, self.last = Cc, H, W, esz, first, first
        self.offset, self.nbytes, self.pinned = -1, 0, False


class LayerPlan:
    """Symbolic pass -> arena assignment -> dy_program."""

    def __init__(self, model, mb: int, H: int, W: int, device, images: torch.Tensor, y: torch.Tensor, fuse_decode: bool = False,
                 head_lanes: Optional[int] = None):
        if H % 32 or W % 32:
            raise _C.DroneYoloError(f"input {H}x{W} must be a multiple of the maximum stride 32")
        self.model, self.mb, self.H, self.W, self.device = model, mb, H, W, device
        self.bufs: list[_Buf] = []
        self.ops: list[dict] = []
        self.keep = []                      # packed weights etc. that must outlive the program
        self.fuse_upsample = not os.environ.get("DY_NO_FUSE_UPSAMPLE")
        self.fuse_tail = not os.environ.get("DY_NO_FUSE_TAIL")
        # decode inside the Detect tails: the raw maps of those levels are never materialised (Engine.raw_maps() is then
        # unavailable); needs the whole batch in one replay because the prediction tensor's address is baked into tensor maps
        self.fuse_cv1 = not os.environ.get("DY_NO_FUSE_CV1")
        # C2f.cv1 behind Concat([Upsample(a), b]) split across the upsample (a 1x1 conv commutes with nearest upsampling):
        # W_a * a at low resolution, then act(W_b * b + bias + up(W_a * a)); the upsampled tensor is never materialised
        self.split_up = not os.environ.get("DY_NO_SPLIT_UP")
        self.up_split: dict[int, tuple] = {}           # concat layer -> (low-resolution source Ref, channels of the upsampled part)
        # side lanes for the Detect branches (dy_program_set_lane): 0 = everything on the caller's stream in layer order,
        # 1 = the branches of every level on one side stream, 2 = box branch and class branch on two side streams.  A level's
        # branches are emitted as soon as its source layer is done, so they overlap the rest of the neck.  Measured on the
        # power-capped s@640 B=64 step (profiles/r02_summary.md): 17.55 k images/s with 0, 1 or 2 lanes alike (the persistent
        # kernels hold one CTA per SM, so a second lane only fills partial waves) -> off by default.
        self.head_lanes = int(os.environ.get("DY_HEAD_LANES", "0")) if head_lanes is None else int(head_lanes)
        self.lane = 0
        self.fuse_decode = (bool(fuse_decode) and self.fuse_tail and not os.environ.get("DY_NO_FUSE_DECODE")
                            and y is not None and mb == y.shape[0])
        self._build_symbolic()
        self._assign_arena()
        self._emit(images, y)

    # ---------------------------------------------------------------------------------------------
    def _producer_of(self, ref):
        """Index of the conv op that writes exactly `ref` (and can take a fused upsample store), else None."""
        for idx in range(len(self.ops) - 1, -1, -1):
            op = self.ops[idx]
            if op.get("out") == ref:
                ok = op["kind"] == "conv" and "up" not in op and ref.c % 8 == 0
                return idx if ok else None
        return None

    def _new_buf(self, Cc, H, W, esz=2):
        self.bufs.append(_Buf(Cc, H, W, esz, len(self.ops)))
        return len(self.bufs) - 1

    def _touch(self, *refs):
        t = len(self.ops)
        for r in refs:
            if r is not None:
                self.bufs[r.buf].last = max(self.bufs[r.buf].last, t)

    def _op(self, **kw):
        self._touch(kw.get("inp"), kw.get("out"), kw.get("res"), kw.get("pre"), kw["tail"][3] if kw.get("tail") else None)
        for r in kw.get("levels", ()):
            self._touch(r)
        kw["lane"] = self.lane
        self.ops.append(kw)

    def _sync(self, waiter, signaller):
        self.ops.append({"kind": "sync", "waiter": waiter, "signaller": signaller, "lane": 0})

    def _build_symbolic(self):
        layers = list(self.model.model)
        n = len(layers)
        # homes: a layer output that feeds a Concat is produced directly inside that Concat's buffer
        home: dict[int, tuple[int, int]] = {}          # producer layer index -> (concat layer index, channel offset)
        cat_ch: dict[int, int] = {}
        out_ch: list[int] = []
        for i, m in enumerate(layers):
            if isinstance(m, Concat):
                srcs = [(j if j >= 0 else i + j) for j in m.f]
                c0 = 0
                for j in srcs:
                    if j in home or isinstance(layers[j], Concat):
                        raise _C.DroneYoloError(f"layer {j} feeds two Concats: plan would need a copy op (not in Drone-YOLO graphs)")
                    home[j] = (i, c0)
                    c0 += out_ch[j]
                cat_ch[i] = c0
                out_ch.append(c0)
            elif isinstance(m, Detect):
                out_ch.append(0)
            elif isinstance(m, Upsample):
                out_ch.append(out_ch[i - 1 if m.f == -1 else m.f])
            else:
                out_ch.append(self._module_cout(m))
        self.layer_ref: list[Optional[Ref]] = [None] * n
        cat_buf: dict[int, int] = {}
        res = {"H": self.H, "W": self.W}
        det = layers[-1] if isinstance(layers[-1], Detect) else None
        det_src = [(j if j >= 0 else n - 1 + j) for j in det.f] if det is not None else []
        self._det_state = None

        def dest(i, c, H, W):
            """Where layer i writes its output."""
            if i in home:
                ci, c0 = home[i]
                if ci not in cat_buf:
                    cat_buf[ci] = self._new_buf(cat_ch[ci], H, W)
                b = self.bufs[cat_buf[ci]]
                if (b.H, b.W) != (H, W):
                    raise _C.DroneYoloError(f"Concat {ci}: resolution mismatch")
                return Ref(cat_buf[ci], c0, c, H, W)
            return Ref(self._new_buf(c, H, W), 0, c, H, W)

        for i, m in enumerate(layers):
            f = m.f
            src = None
            if isinstance(f, int):
                src = self.layer_ref[i - 1 if f == -1 else f] if i > 0 else None
            if isinstance(m, Conv) and m.is_stem:
                Ho, Wo = self.H // 2, self.W // 2
                out = dest(i, m.conv.out_channels, Ho, Wo)
                self._op(kind="stem", mod=m, out=out)
            elif isinstance(m, DWConv):
                out = dest(i, m.conv.out_channels, src.H // 2, src.W // 2)
                self._op(kind="dwconv", mod=m, inp=src, out=out)
            elif isinstance(m, (Conv, RepVGGBlock, RepConv)):
                cout, k, s = self._conv_geom(m)
                out = dest(i, cout, src.H // s, src.W // s)
                self._op(kind="conv", mod=m, inp=src, out=out)
            elif isinstance(m, C2f):
                out = dest(i, m.cv2.conv.out_channels, src.H, src.W)
                # the layer in front feeds only this C2f: its output can stay on the SM (see _emit_c2f)
                j = i - 1 if f == -1 else f
                private = (self.fuse_cv1 and isinstance(f, int) and j not in home and j not in getattr(self.model, "save", ())
                           and self.ops and self.ops[-1]["kind"] == "conv" and self.ops[-1].get("out") == src)
                self._emit_c2f(m, src, out, len(self.ops) - 1 if private else None, split=self.up_split.get(j))
            elif isinstance(m, SPPF):
                out = dest(i, m.cv2.conv.out_channels, src.H, src.W)
                c_ = m.cv1.conv.out_channels
                cat = self._new_buf(4 * c_, src.H, src.W)
                self._op(kind="conv", mod=m.cv1, inp=src, out=Ref(cat, 0, c_, src.H, src.W))
                self._op(kind="pool", inp=Ref(cat, 0, 4 * c_, src.H, src.W), c=c_)
                self._op(kind="conv", mod=m.cv2, inp=Ref(cat, 0, 4 * c_, src.H, src.W), out=out)
            elif isinstance(m, Upsample):
                out = dest(i, src.c, src.H * 2, src.W * 2)
                prod = self._producer_of(src)
                ci = home[i][0] if i in home else None
                nxt = layers[ci + 1] if ci is not None and ci + 1 < n else None
                users = [li for li, l in enumerate(layers) for f_ in (l.f if isinstance(l.f, (list, tuple)) else [l.f])
                         if (li + f_ if f_ < 0 else f_) == ci] if ci is not None else []
                if (self.split_up and isinstance(nxt, C2f) and users == [ci + 1] and (home[i][1] == 0 or home[i][1] + src.c == cat_ch[ci])
                        and ci not in getattr(self.model, "save", ()) and nxt.cv1.conv.kernel_size[0] == 1
                        and isinstance(nxt.cv1.act, nn.SiLU) and src.c % 8 == 0 and ((2 * nxt.c) % 32 == 0 or 2 * nxt.c in (8, 16, 24, 32))
                        and src.H * 2 == out.H and src.W * 2 == out.W):
                    # the Concat only feeds that C2f and the upsampled part sits first or last in it (the -sf graphs put it last):
                    # nothing is written here
                    self.up_split[ci] = (src, src.c, home[i][1])
                elif prod is not None and self.fuse_upsample:
                    # nn.Upsample(2x nearest) folded into its producer: the conv's epilogue also stores every output tile
                    # into the four (dy, dx) parity views of this destination (dy_conv_desc.up_out)
                    self.ops[prod]["up"] = out
                    b = self.bufs[out.buf]
                    b.first = min(b.first, prod)
                    b.last = max(b.last, prod)
                else:
                    self._op(kind="upsample", inp=src, out=out)
            elif isinstance(m, Concat):
                b = self.bufs[cat_buf[i]]
                out = Ref(cat_buf[i], 0, b.C, b.H, b.W)
            elif isinstance(m, Detect):
                for li, j in enumerate(det_src):             # levels not emitted yet (side lanes off)
                    if self._det_state is None or li not in self._det_state["done"]:
                        self._emit_detect_level(m, li, self.layer_ref[j])
                self._finish_detect(m)
                out = None
            else:
                raise _C.DroneYoloError(f"layer {i}: {type(m).__name__} has no plan lowering")
            self.layer_ref[i] = out
            if self.head_lanes > 0 and det is not None and i in det_src and out is not None:
                # this pyramid level is complete: its Detect branches only depend on it (head.py:64-74) -> side lane(s), now
                self._emit_detect_level(det, det_src.index(i), out)
        del res

    @staticmethod
    def _module_cout(m):
        if isinstance(m, (Conv, DWConv)):
            return m.conv.out_channels
        if isinstance(m, (RepVGGBlock, RepConv)):
            return m._geom()[0]
        if isinstance(m, (C2f, SPPF)):
            return m.cv2.conv.out_channels
        raise _C.DroneYoloError(f"{type(m).__name__}: unknown output width")

    @staticmethod
    def _conv_geom(m):
        if isinstance(m, Conv):
            return m.conv.out_channels, m.conv.kernel_size[0], m.conv.stride[0]
        cout, s = m._geom()
        return cout, 3, s

    def _emit_c2f(self, m: C2f, src: Ref, out: Ref, prod: Optional[int] = None, split=None):
        c, n = m.c, len(m.m)
        H, W = src.H, src.W
        cat = self._new_buf((2 + n) * c, H, W)
        fused = False
        if split is not None:
            lo, ca, c0 = split                               # low-resolution source of the upsampled part, its channels and offset
            w, b = m.cv1.fused_weight_bias()                 # fp32 [2c, C, 1, 1], [2c]; the upsampled part is channels [c0, c0 + ca)
            wa, _ = K.pack_conv_weight(w[:, c0:c0 + ca].contiguous(), None)
            sk0 = ca if c0 == 0 else 0                       # the skip tensors: the rest of the Concat, before or behind it
            wb, bb = K.pack_conv_weight(w[:, sk0:sk0 + src.c - ca].contiguous(), b)
            zero = torch.zeros_like(bb)
            t = Ref(self._new_buf(2 * c, lo.H, lo.W, esz=4), 0, 2 * c, lo.H, lo.W)
            self._op(kind="conv", w=(wa, zero), cout=2 * c, k=1, s=1, act=False, inp=lo, out=t)
            self._op(kind="conv", w=(wb, bb), cout=2 * c, k=1, s=1, act=True, inp=Ref(src.buf, src.c0 + sk0, src.c - ca, H, W),
                     out=Ref(cat, 0, 2 * c, H, W), pre=t)
            fused = True
        if prod is not None:
            # cv1 (1x1 Conv + SiLU, block.py:236) as the fused tail of the 3x3 stride-2 conv in front of it (the RepVGG
            # downsample at P2 of the s scale: 32 -> 64, then 64 -> 64): the 64-channel tensor between them, 3.3 MB per
            # image written and read back, never reaches HBM (dy_conv_desc.tail_decode == 3)
            po = self.ops[prod]
            cout, k, s_ = self._conv_geom(po["mod"]) if "mod" in po else (po["cout"], po["k"], po["s"])
            act = isinstance(getattr(po.get("mod"), "act", getattr(po.get("mod"), "nonlinearity", None)), nn.SiLU) if "mod" in po else po["act"]
            if (k == 3 and s_ == 2 and po["inp"].c <= 32 and cout == 64 and 2 * c <= 64 and (2 * c) % 8 == 0 and act
                    and not any(key in po for key in ("up", "res", "tail")) and m.cv1.conv.kernel_size[0] == 1
                    and isinstance(m.cv1.act, nn.SiLU)):
                w1, b1 = m.cv1.packed()
                po["tail"] = (w1, b1, 2 * c, Ref(cat, 0, 2 * c, H, W))
                po["out"] = None
                self.bufs[src.buf].C = 0                   # never materialised
                cb = self.bufs[cat]
                cb.first = min(cb.first, prod)
                fused = True
        if not fused:
            self._op(kind="conv", mod=m.cv1, inp=src, out=Ref(cat, 0, 2 * c, H, W))
        for i, b in enumerate(m.m):
            xin = Ref(cat, (1 + i) * c, c, H, W)
            tmp = Ref(self._new_buf(b.cv1.conv.out_channels, H, W), 0, b.cv1.conv.out_channels, H, W)
            self._op(kind="conv", mod=b.cv1, inp=xin, out=tmp)
            self._op(kind="conv", mod=b.cv2, inp=tmp, out=Ref(cat, (2 + i) * c, c, H, W), res=xin if b.add else None)
        self._op(kind="conv", mod=m.cv2, inp=Ref(cat, 0, (2 + n) * c, H, W), out=out)

    def _emit_detect_level(self, m: Detect, i: int, src: Ref):
        """The three convs of pyramid level `i` (head.py:41-47, 64-74): merged first 3x3 (N = c2 + c3), then per branch
        Conv(c,c,3) -> nn.Conv2d(c,n,1), fused (+ decode) where the 64-channel halo kernel applies."""
        if self._det_state is None:
            strides = [int(v) for v in m.stride.tolist()]
            hw = [(self.H // st) * (self.W // st) for st in strides]
            self._det_state = {"packed": m.packed(), "done": set(), "levels": {}, "all": {}, "forks": {},
                               "a_off": [sum(hw[:k]) for k in range(len(hw))]}
        st = self._det_state
        packed, a_off = st["packed"], st["a_off"][i]
        lanes = self.head_lanes
        if lanes > 0:
            st["forks"][i] = len(self.ops)
            self._sync(1, 0)                        # lane 1 continues from here: the level's source is complete
            self.lane = 1
        H, W = src.H, src.W
        if H * W != (self.H // int(m.stride[i])) * (self.W // int(m.stride[i])):
            raise _C.DroneYoloError(f"Detect level {i}: source resolution {H}x{W} does not match stride {int(m.stride[i])}")
        c2, c3 = m.cv2[i][0].conv.out_channels, m.cv3[i][0].conv.out_channels
        (wf, bf), (wbx, bbx), (wcl, bcl) = packed[i]
        t1 = self._new_buf(c2 + c3, H, W)
        ncp = m.raw_ld - 4 * m.reg_max          # class logits padded to 16 channels (zero weights) -> TMA-store path
        self._op(kind="conv", w=(wf, bf), cout=c2 + c3, k=3, s=1, act=True, inp=src, out=Ref(t1, 0, c2 + c3, H, W))
        # Each branch ends Conv(c,c,3) -> nn.Conv2d(c,n,1) (head.py:41-47).  Where the 64-channel halo kernel applies, the
        # 1x1 runs inside the 3x3's epilogue (dy_conv_desc.weight2) and the intermediate tensor is never written; with
        # fuse_decode the same epilogue also decodes the logits (dy_conv_desc.tail_decode) and the raw map is skipped too.
        # ragged 8x16 tiles waste MMA rows, but a level this small is launch-bound: three launches less win
        halo_ok = self.fuse_tail and (H * W / (-(-W // 8) * 8 * -(-H // 16) * 16) >= 0.8 or self.mb * H * W <= 65536)
        branches = ((m.cv2[i][1], 0, c2, (wbx, bbx), 4 * m.reg_max, 0), (m.cv3[i][1], c2, c3, (wcl, bcl), ncp, 4 * m.reg_max))
        fusable = [halo_ok and cw == 64 and cout1 <= 64 for (_, _, cw, _, cout1, _) in branches]
        dec = self.fuse_decode and all(fusable) and m.nc <= 32 and W % 4 == 0 and a_off % 4 == 0
        raw = None if dec else self._new_buf(m.raw_ld, H, W, esz=4)
        t2 = None
        for bi, (mod3, cin0, cw, (w1, b1), cout1, c0out) in enumerate(branches):
            if lanes >= 2 and bi == 1:
                self._sync(2, 1)                    # the class branch runs beside the box branch, behind the merged first conv
                self.lane = 2
            if fusable[bi]:
                tail = ((w1, b1, cout1, None, (bi + 1, a_off, float(m.stride[i]))) if dec
                        else (w1, b1, cout1, Ref(raw, c0out, cout1, H, W)))
                self._op(kind="conv", mod=mod3, inp=Ref(t1, cin0, cw, H, W), out=None, tail=tail)
            else:
                if t2 is None:
                    t2 = self._new_buf(c2 + c3, H, W)
                self._op(kind="conv", mod=mod3, inp=Ref(t1, cin0, cw, H, W), out=Ref(t2, cin0, cw, H, W))
                self._op(kind="conv", w=(w1, b1), cout=cout1, k=1, s=1, act=False, inp=Ref(t2, cin0, cw, H, W),
                         out=Ref(raw, c0out, cout1, H, W))
        self.lane = 0
        if not dec:
            st["levels"][i] = (Ref(raw, 0, m.no, H, W), a_off, float(m.stride[i]))
        st["all"][i] = None if dec else Ref(raw, 0, m.no, H, W)
        st["done"].add(i)

    def _finish_detect(self, m: Detect):
        st = self._det_state
        for lane in range(1, min(self.head_lanes, 2) + 1):
            self._sync(0, lane)                     # join: the decode of the remaining levels and the NMS read every branch's output
        if self.head_lanes > 0:
            # Side-lane ops run concurrently with the main chain: whatever they touch is live from the fork to the end of the
            # program (the arena otherwise recycles a buffer after its last reader IN PROGRAM ORDER)
            end = len(self.ops)
            fork_of = {}
            cur = None
            for idx, op in enumerate(self.ops):
                if op["kind"] == "sync" and op["waiter"] == 1 and op["signaller"] == 0:
                    cur = idx
                if op.get("lane", 0) > 0 and op["kind"] != "sync":
                    fork_of[idx] = cur
            for idx, fork in fork_of.items():
                op = self.ops[idx]
                refs = [op.get("inp"), op.get("out"), op.get("res"), op["tail"][3] if op.get("tail") else None]
                for r in refs:
                    if r is not None:
                        b = self.bufs[r.buf]
                        b.first = min(b.first, fork)
                        b.last = max(b.last, end)
        order = sorted(st["levels"])
        if order:
            self._op(kind="decode", levels=[st["levels"][i][0] for i in order], det=m, strides=[st["levels"][i][2] for i in order],
                     anchor_off=[st["levels"][i][1] for i in order] if len(order) < len(st["all"]) else None)
        self.raw_refs = [st["all"][i] for i in sorted(st["all"])]

    # ---------------------------------------------------------------------------------------------
    def _assign_arena(self, allocate: bool = True):
        """First-fit offsets with lifetime reuse (buffers are free after their last reader)."""
        live: list[tuple[int, int, int]] = []      # (offset, nbytes, last)
        total = 0
        order = sorted(range(len(self.bufs)), key=lambda i: self.bufs[i].first)
        for bi in order:
            b = self.bufs[bi]
            b.nbytes = (self.mb * b.H * b.W * b.C * b.esz + 1023) // 1024 * 1024
            live = [x for x in live if x[2] >= b.first]
            live.sort()
            off = 0
            for o, nb, _ in live:
                if off + b.nbytes <= o:
                    break
                off = max(off, o + nb)
            b.offset = off
            live.append((off, b.nbytes, b.last))
            total = max(total, off + b.nbytes)
        self.arena_bytes = total
        if allocate:
            self.arena = torch.empty((max(total, 1024),), device=self.device, dtype=torch.uint8)

    def tensor(self, r: Ref) -> torch.Tensor:
        """(mb, c, H, W) channels-last view of a Ref inside the arena."""
        b = self.bufs[r.buf]
        dt = torch.float32 if b.esz == 4 else torch.bfloat16
        flat = self.arena[b.offset: b.offset + self.mb * b.H * b.W * b.C * b.esz].view(dt)
        return flat.view(self.mb, b.H, b.W, b.C).permute(0, 3, 1, 2)[:, r.c0: r.c0 + r.c]

    # ---------------------------------------------------------------------------------------------
    def _emit(self, images: torch.Tensor, y: torch.Tensor):
        lib = _C.lib()
        h = C.c_void_p()
        _C.check(lib.dy_program_create(C.byref(h)), "dy_program_create")
        self.handle = h
        mb = self.mb
        lane = 0
        for op in self.ops:
            kind = op["kind"]
            if kind == "sync":
                for ln in (op["waiter"], op["signaller"]):          # a side lane exists once it has been selected
                    if ln > 0:
                        _C.check(lib.dy_program_set_lane(h, ln), "set_lane")
                _C.check(lib.dy_program_set_lane(h, lane), "set_lane")
                _C.check(lib.dy_program_add_sync(h, op["waiter"], op["signaller"]), "add_sync")
                continue
            if op.get("lane", 0) != lane:
                lane = op.get("lane", 0)
                _C.check(lib.dy_program_set_lane(h, lane), "set_lane")
            if kind == "stem":
                w, b = op["mod"].packed()
                self.keep.append((w, b))
                out = self.tensor(op["out"])
                optr, old, *_ = K.nhwc_view(out)
                in_dt = _C.DY_U8 if images.dtype == torch.uint8 else _C.DY_F32
                _C.check(lib.dy_program_add_stem(h, images.data_ptr(), in_dt, mb, self.H, self.W, w.data_ptr(), b.data_ptr(),
                                                 w.shape[0], optr, old), "add_stem")
            elif kind == "conv":
                if "mod" in op:
                    w, b = op["mod"].packed()
                    cout, k, s = self._conv_geom(op["mod"])
                    act = isinstance(getattr(op["mod"], "act", getattr(op["mod"], "nonlinearity", None)), nn.SiLU)
                else:
                    (w, b), cout, k, s, act = op["w"], op["cout"], op["k"], op["s"], op["act"]
                self.keep.append((w, b))
                res = self.tensor(op["res"]) if op.get("res") is not None else None
                up = self.tensor(op["up"]) if op.get("up") is not None else None
                tail = None
                if op.get("tail") is not None:
                    w2, b2, cout2, r2 = op["tail"][:4]
                    self.keep.append((w2, b2))
                    if r2 is None:                  # fused decode: (mode, first anchor, stride) -> the engine's prediction tensor
                        mode, a0, st = op["tail"][4]
                        tail = (w2, b2, cout2, None, (mode, y, a0, st))
                    else:
                        tail = (w2, b2, cout2, self.tensor(r2))
                out_t = self.tensor(op["out"]) if op.get("out") is not None else None
                pre = self.tensor(op["pre"]) if op.get("pre") is not None else None
                d = K.conv_desc(self.tensor(op["inp"]), w, b, cout, k, s, act, out_t, res, up, tail, pre)
                _C.check(lib.dy_program_add_conv(h, C.byref(d)), "add_conv")
            elif kind == "pool":
                t = self.tensor(op["inp"])
                p, ld, B, H, W, _ = K.nhwc_view(t)
                _C.check(lib.dy_program_add_sppf_pool(h, p, B, H, W, op["c"], ld), "add_sppf_pool")
            elif kind == "upsample":
                ti, to = self.tensor(op["inp"]), self.tensor(op["out"])
                ip, ild, B, H, W, Cc = K.nhwc_view(ti)
                op_, old, *_ = K.nhwc_view(to)
                _C.check(lib.dy_program_add_upsample2x(h, ip, ild, B, H, W, Cc, op_, old), "add_upsample2x")
            elif kind == "dwconv":
                w, b = op["mod"].packed()
                w = w.reshape(w.shape[0], 18).contiguous()
                self.keep.append((w, b))
                ti, to = self.tensor(op["inp"]), self.tensor(op["out"])
                ip, ild, B, H, W, Cin = K.nhwc_view(ti)
                op_, old, *_ = K.nhwc_view(to)
                _C.check(lib.dy_program_add_dwconv3x3s2(h, ip, ild, B, H, W, Cin, w.data_ptr(), b.data_ptr(), w.shape[0],
                                                        op_, old), "add_dwconv")
            elif kind == "decode":
                det = op["det"]
                levels = [self.tensor(r) for r in op["levels"]]
                d = K.decode_desc(levels, op["strides"], det.nc, y, op.get("anchor_off"))
                d.B = mb
                _C.check(lib.dy_program_add_decode(h, C.byref(d)), "add_decode")
            else:
                raise AssertionError(kind)
        self.launches = lib.dy_program_num_launches(h)

    def describe(self) -> list[str]:
        """One line per op, in program order (what `profile` times)."""
        names = []
        for op in self.ops:
            kind = op["kind"]
            if kind == "sync":
                names.append(f"sync lane {op['waiter']} <- lane {op['signaller']}")
            elif kind == "conv":
                if "mod" in op:
                    cout, k, s = self._conv_geom(op["mod"])
                else:
                    cout, k, s = op["cout"], op["k"], op["s"]
                inp, tail = op["inp"], op.get("tail")
                txt = f"conv{k}x{k}s{s} {inp.c}->{cout} @{inp.H // s}x{inp.W // s}"
                if op.get("res") is not None:
                    txt += " +res"
                if op.get("up") is not None:
                    txt += " +up2x"
                if op.get("pre") is not None:
                    txt += " +pre(up)"
                if tail:
                    txt += f" +1x1->{tail[2]}" + (" +decode" if len(tail) > 4 else "")
                names.append(txt)
            elif kind == "stem":
                names.append(f"stem 3->{op['out'].c} @{op['out'].H}x{op['out'].W}")
            elif kind == "pool":
                names.append(f"sppf pool c={op['c']} @{op['inp'].H}x{op['inp'].W}")
            elif kind == "decode":
                names.append(f"detect decode ({len(op['levels'])} levels)")
            else:
                names.append(kind)
        return names

    def profile(self, in_offset_bytes: int = 0, out_offset_bytes: int = 0, reps: int = 5, stream: Optional[int] = None):
        """Per-op device milliseconds of one eager replay (`dy_program_profile`): the role of the reference's per-layer profile
        (`BaseModel._profile_one_layer`, nn/tasks.py:171-191).  Returns [(description, ms)], sync ops left out."""
        lib = _C.lib()
        n = lib.dy_program_num_ops(self.handle)
        names = self.describe()
        if n != len(names):
            raise _C.DroneYoloError(f"plan holds {len(names)} ops, the program {n}")
        ms = (C.c_float * n)()
        with torch.cuda.device(self.device):
            s = _C.stream_ptr(self.device) if stream is None else stream
            _C.check(lib.dy_program_profile(self.handle, in_offset_bytes, out_offset_bytes, s, int(reps), ms, n), "dy_program_profile")
        return [(nm, float(t)) for nm, t, op in zip(names, ms, self.ops) if op["kind"] != "sync"]

    def run(self, in_offset_bytes: int, out_offset_bytes: int, stream: int):
        _C.check(_C.lib().dy_program_run(self.handle, in_offset_bytes, out_offset_bytes, stream), "dy_program_run")

    def raw_maps(self):
        """The raw head maps of the LAST replayed micro-batch, as the reference's (mb, no, H, W) views."""
        if any(r is None for r in self.raw_refs):
            raise _C.DroneYoloError("raw head maps are not materialised when the Detect decode is fused into the conv tails "
                                    "(build the Engine with fuse_decode=False to inspect them)")
        return [self.tensor(r) for r in self.raw_refs]

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _C.lib().dy_program_destroy(self.handle)
        except Exception:  # noqa: BLE001
            pass
