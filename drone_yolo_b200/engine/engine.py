"""Engine: the device-side runtime of one (model, batch, H, W) configuration.

Plays the role of the reference's AutoBackend (ultralytics/nn/autobackend.py:149-159,548-556: fuse, dtype, forward)
plus DetectionPredictor.postprocess's NMS call (models/yolo/detect/predict.py:25-35) for the PyTorch branch only.
Owns: the static input batch, the layer plan (engine/plan.py), the (B, 4+nc, A) prediction, the NMS buffers and
one CUDA graph that replays [plan x micro-batches, NMS].
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from .. import _C
from .. import kernels as K
from .plan import LayerPlan


def pick_micro_batch(model, batch: int, H: int, W: int, cap: int = 64) -> int:
    """Images per replay of the layer plan.  Measured on B200 (profiles/r01_summary.md): while the small layers are
    still launch/latency-bound the largest micro-batch wins (s@640: 2 -> 2.0k, 8 -> 4.7k, 32 -> 6.6k, 64 -> 7.0k img/s
    with the first kernels), so the default is the largest divisor of `batch` up to `cap`; smaller values keep a
    layer's output L2-resident (126 MB) for its consumer and become interesting once the kernels are memory-bound."""
    best = 1
    for d in range(1, min(batch, cap) + 1):
        if batch % d == 0:
            best = d
    return best


class Engine:
    def __init__(self, model, batch: int, imgsz, device, micro_batch: int = 0, conf: float = 0.25, iou: float = 0.7,
                 max_det: int = 300, classes=None, agnostic: bool = False, multi_label: bool = False,
                 max_nms: int = 30000, max_wh: float = 7680.0, cuda_graph: bool = True,
                 input_dtype: torch.dtype = torch.float32, fuse_decode: bool = True, input_slots: int = 1,
                 head_lanes: Optional[int] = None, rescale: bool = False):
        device = torch.device(device)
        if device.type != "cuda":
            raise _C.DroneYoloError("drone_yolo_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        _C.check(_C.lib().dy_device_check(device.index or 0), "dy_device_check")
        H, W = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        self.model, self.device, self.batch, self.H, self.W = model, device, batch, H, W
        det = model.model[-1]
        self.nc = det.nc
        self.A = sum((H // int(s)) * (W // int(s)) for s in det.stride.tolist())
        self.mb = micro_batch if micro_batch and batch % micro_batch == 0 else pick_micro_batch(model, batch, H, W)
        with torch.cuda.device(device):
            if input_dtype not in (torch.float32, torch.uint8):
                raise _C.DroneYoloError("engine input must be float32 in [0,1] or uint8 0..255")
            # `input_slots` > 1: several resident input batches in one allocation, so that the upload of batch i+1 can land
            # while batch i computes and `step(slot)` reads it in place (no device-side hand-over copy); one graph per slot
            self.input_slots = max(1, int(input_slots))
            self._all_images = torch.zeros((self.input_slots * batch, 3, H, W), device=device, dtype=input_dtype)
            self.image_slots = [self._all_images[k * batch:(k + 1) * batch] for k in range(self.input_slots)]
            self.images = self.image_slots[0]
            self.y = torch.empty((batch, 4 + self.nc, self.A), device=device, dtype=torch.float32)
            self.plan = LayerPlan(model, self.mb, H, W, device, self._all_images, self.y, fuse_decode=fuse_decode,
                                  head_lanes=head_lanes)
            ml = bool(multi_label) and self.nc > 1
            self.nms_bufs = K.NmsBuffers(batch, self.nc, self.A, max_det, ml, device)
            classes = K.normalize_classes(classes)
            self.nms_cfg = dict(conf=conf, iou=iou, max_det=max_det, classes=classes, agnostic=agnostic, multi_label=ml,
                                max_nms=max_nms, max_wh=max_wh)
            # `rescale`: the NMS output phase also maps the rows to each original image (scale_boxes + clip_boxes, ops.py:92-127,
            # 335-354); `self.rescale` holds (pad_x, pad_y, gain, w0, h0) per image, identity until the predictor fills it
            # (one parameter block per input slot: the upload of batch i+1's block must not race the NMS of batch i)
            self.rescale_slots = None
            if rescale:
                self.rescale_slots = torch.zeros((self.input_slots, batch, 8), device=device, dtype=torch.float32)
                self.rescale_slots[..., 2] = 1.0
                self.rescale_slots[..., 3] = float(W)
                self.rescale_slots[..., 4] = float(H)
            self._nms_progs = []
            for k in range(self.input_slots if rescale else 1):
                d = K.nms_desc(self.y, self.nms_bufs, conf, iou, max_det, max_nms, max_wh, agnostic, ml, classes,
                               in_place=False, want_kept=True, rescale=self.rescale_slots[k] if rescale else None)
                h = C.c_void_p()
                _C.check(_C.lib().dy_program_create(C.byref(h)), "dy_program_create")
                self._nms_progs.append(h)
                _C.check(_C.lib().dy_program_add_nms(h, C.byref(d)), "dy_program_add_nms")
            self._nms_prog = self._nms_progs[0]
            self.launches_per_step = self.plan.launches * (batch // self.mb) + _C.lib().dy_program_num_launches(self._nms_prog)
            self.graph: Optional[torch.cuda.CUDAGraph] = None
            self.graphs: list = []
            self.enqueue()                       # eager warm-up (sets kernel attributes, pages in code)
            torch.cuda.synchronize(device)
            if cuda_graph:
                # captured on a prioritised stream: the plan's side lanes (Detect branches, lowest priority) then yield the
                # SMs to the main chain whenever both have CTAs waiting
                cap = torch.cuda.Stream(device=device, priority=-1)
                for k in range(self.input_slots):
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=cap):
                        self.enqueue(slot=k)
                    self.graphs.append(g)
                self.graph = self.graphs[0]

    def enqueue(self, stream: Optional[int] = None, nms: bool = True, slot: int = 0):
        """Enqueue the whole step on `stream` (default: torch's current stream), reading input slot `slot`."""
        with torch.cuda.device(self.device):       # the library launches on the calling thread's current device
            s = _C.stream_ptr(self.device) if stream is None else stream
            in_bytes = self.mb * 3 * self.H * self.W * self.images.element_size()
            out_bytes = self.mb * (4 + self.nc) * self.A * 4
            slot_bytes = self.batch * 3 * self.H * self.W * self.images.element_size()
            for m in range(self.batch // self.mb):
                self.plan.run(slot * slot_bytes + m * in_bytes, m * out_bytes, s)
            if nms:
                _C.check(_C.lib().dy_program_run(self._nms_progs[slot % len(self._nms_progs)], 0, 0, s), "dy_program_run(nms)")

    def step(self, slot: int = 0):
        """Run conv stack + decode + NMS on the resident batch of input slot `slot`; results in y / nms_bufs (no host sync)."""
        if self.graphs:
            if self.device.index != torch.cuda.current_device():
                with torch.cuda.device(self.device):
                    self.graphs[slot].replay()
            else:
                self.graphs[slot].replay()
        else:
            self.enqueue(slot=slot)
        return self.nms_bufs.out, self.nms_bufs.counts

    def __call__(self, images: Optional[torch.Tensor] = None):
        if images is not None:
            if images.shape != self.images.shape:
                raise _C.DroneYoloError(f"engine built for {tuple(self.images.shape)}, got {tuple(images.shape)}")
            self.images.copy_(images, non_blocking=True)
        return self.step()

    def profile(self, reps: int = 5, verbose: bool = False):
        """Per-layer device times of the resident batch (the reference's `predict(profile=True)`, nn/tasks.py:116,151-152,171-191):
        [(op, ms)] for every op of the conv-stack plan (first micro-batch) and the NMS program."""
        rows = self.plan.profile(0, 0, reps)
        lib = _C.lib()
        n = lib.dy_program_num_ops(self._nms_prog)
        ms = (C.c_float * n)()
        with torch.cuda.device(self.device):
            _C.check(lib.dy_program_profile(self._nms_prog, 0, 0, _C.stream_ptr(self.device), int(reps), ms, n), "dy_program_profile(nms)")
        rows += [("nms (filter + select)", float(ms[i])) for i in range(n)]
        if verbose:
            for nm, t in rows:
                print(f"{t * 1e3:9.1f} us  {nm}")
            print(f"{sum(t for _, t in rows) * 1e3:9.1f} us  total ({len(rows)} ops, micro-batch {self.mb})")
        return rows

    def raw_maps(self):
        return self.plan.raw_maps()

    def __del__(self):
        try:
            for h in getattr(self, "_nms_progs", []):
                _C.lib().dy_program_destroy(h)
        except Exception:  # noqa: BLE001
            pass
