"""DetectionPredictor: source -> preprocess -> Engine (conv stack + decode + NMS) -> Results.

Mirror of the reference's BasePredictor / DetectionPredictor (ultralytics/engine/predictor.py:66-410,
models/yolo/detect/predict.py:23-73) for the inference hot path: same constructor (`overrides`, `_callbacks`),
`setup_model`, `__call__(source, stream)`, `preprocess`, `inference`, `postprocess`, the five `on_predict_*`
callback events and `Results.speed`.  Differences:
  * the batch loop (predictor.py:251-288) is a two-slot pipeline: batch i+1 is staged and uploaded (copy stream) while
    batch i computes and batch i-1 is turned into Results on the host; no device-wide synchronisation, `Results.speed`
    comes from CUDA events;
  * one D2H of the padded detections per batch instead of per-image boolean-mask syncs; scale_boxes + clip_boxes run in
    the NMS output phase (dy_nms_desc.rescale); `orig_img` of tensor sources is produced lazily;
  * under torch.distributed (one process per GPU) every batch is sharded over the ranks and gathered to rank 0, which
    builds the Results (the reference's `select_device("0,1")`, utils/torch_utils.py:202-219, silently uses cuda:0).
"""
from __future__ import annotations

import threading
import time
from collections import OrderedDict, deque
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from .. import _C
from .. import kernels as K
from ..parallel import DetectionGather, shard_bounds, shard_of
from ..utils import ops
from .engine import Engine
from .results import Results

DEFAULTS = dict(task="detect", mode="predict", imgsz=640, batch=1, device=None, conf=0.25, iou=0.7, max_det=300,
                half=False, classes=None, agnostic_nms=False, augment=False, stream=False, verbose=False,
                micro_batch=0, cuda_graph=True, multi_label=False, gpu_preprocess=True, engine_cache=8, vid_stride=1,
                dtype="bf16", distributed=True)
IMG_FORMATS = {"bmp", "dng", "jpeg", "jpg", "mpo", "png", "tif", "tiff", "webp", "pfm"}                       # data/utils.py:38 (heic needs pillow-heif: not read)
VID_FORMATS = {"asf", "avi", "gif", "m4v", "mkv", "mov", "mp4", "mpeg", "mpg", "ts", "wmv", "webm"}           # data/utils.py:39
EVENTS = ("on_predict_start", "on_predict_batch_start", "on_predict_postprocess_end", "on_predict_batch_end", "on_predict_end")


def letterbox_geometry(shape, new_shape, stride=32, auto=False):
    """(new_w, new_h, left, top, out_h, out_w) of the reference's LetterBox for an image of `shape` (h, w)
    (data/augment.py:1573-1598: ratio, rounded new size, centre padding rounded with -0.1 / +0.1)."""
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


def letterbox(im: np.ndarray, new_shape, stride=32, auto=False, color=(114, 114, 114)):
    """Resize + pad to `new_shape` keeping the aspect ratio (reference data/augment.py:1544-1610, centre padding)."""
    import cv2

    shape = im.shape[:2]
    r = min(new_shape[0] / shape[0], new_shape[1] / shape[1])
    new_unpad = int(round(shape[1] * r)), int(round(shape[0] * r))
    dw, dh = new_shape[1] - new_unpad[0], new_shape[0] - new_unpad[1]
    if auto:
        dw, dh = np.mod(dw, stride), np.mod(dh, stride)
    dw /= 2
    dh /= 2
    if shape[::-1] != new_unpad:
        im = cv2.resize(im, new_unpad, interpolation=cv2.INTER_LINEAR)
    top, bottom = int(round(dh - 0.1)), int(round(dh + 0.1))
    left, right = int(round(dw - 0.1)), int(round(dw + 0.1))
    return cv2.copyMakeBorder(im, top, bottom, left, right, cv2.BORDER_CONSTANT, value=color)


class DetectionPredictor:
    def __init__(self, cfg=None, overrides=None, _callbacks=None):
        args = dict(DEFAULTS)
        args.update(cfg or {})
        args.update(overrides or {})
        if args.get("conf") is None:
            args["conf"] = 0.25
        self.args = SimpleNamespace(**args)
        self._check_args()
        self.model = None
        self.device = None
        self.engines: "OrderedDict" = OrderedDict()              # LRU: key -> Engine
        self.callbacks = {e: [] for e in EVENTS}
        for k, v in (_callbacks or {}).items():
            self.callbacks.setdefault(k, []).extend(v if isinstance(v, (list, tuple)) else [v])
        self.results = None
        self.batch = None
        self.save_dir = None
        self.seen = 0
        self._lock = threading.Lock()
        self._pinned: "OrderedDict" = OrderedDict()              # LRU staging buffers (bounded, see _staging)
        self._copy_stream = None
        self._gather = None

    def _check_args(self):
        """Precision and TTA switches of the reference that this path cannot honour raise instead of being ignored.
        `half` (cfg/default.yaml:54, autobackend.py:132,157): the reference's only precision switch, fp32 vs fp16.  Here the
        conv stack ALWAYS runs bf16 with fp32 accumulation (raw maps within rtol 2e-2 of the fp32 reference) and decode + NMS
        ALWAYS run fp32, so `half=True` and `half=False` are both accepted and select the same arithmetic; an explicit request
        for another precision (`dtype="fp32"` / "fp16") raises."""
        a = self.args
        if getattr(a, "augment", False):
            raise _C.DroneYoloError("augment=True (TTA) is outside the inference hot path")
        if str(getattr(a, "dtype", "bf16")).lower() not in ("bf16", "bfloat16"):
            raise _C.DroneYoloError(f"dtype={a.dtype!r}: the sm_100a conv stack is bf16 (fp32 accumulate; decode and NMS fp32) - there is "
                                    "no fp32 / fp16 conv path; use the reference for those")
        a.classes = K.normalize_classes(getattr(a, "classes", None))   # int / tensor / ndarray / list -> list of ints (mix6.py:80: classes=0)

    def update_args(self, args: dict):
        """Later predict() calls only update the arguments (engine/model.py:555-557); a rejected value is not kept."""
        old = dict(self.args.__dict__)
        self.args.__dict__.update(args)
        try:
            self._check_args()
        except Exception:
            self.args.__dict__.clear()
            self.args.__dict__.update(old)
            raise

    # ---- setup ---------------------------------------------------------------------------------
    def setup_model(self, model, verbose=False):
        """Move to the device, fold BN + re-parameterise (AutoBackend: autobackend.py:149-159)."""
        dev = self.args.device
        if dev is None or dev == "":
            dev = "cuda:0"
        if isinstance(dev, int) or (isinstance(dev, str) and dev.isdigit()):
            dev = f"cuda:{dev}"
        if isinstance(dev, str) and "," in dev:
            # the reference accepts "0,1,.." and then runs on the first device only (torch_utils.py:202-219); multi-GPU here is
            # one process per GPU under torchrun (each rank passes its own device), see _dist()
            raise _C.DroneYoloError(f"device '{dev}': one process drives one GPU - launch with torchrun (one rank per GPU) to shard batches")
        self.device = torch.device(dev)
        if self.device.type != "cuda":
            raise _C.DroneYoloError(f"device '{dev}': drone_yolo_b200 has no CPU path (use the reference for CPU inference)")
        self.model = model.to(self.device).eval().fuse(verbose=False)
        self._flush()
        self.engines.clear()
        self._pinned.clear()
        self._copy_stream = torch.cuda.Stream(device=self.device)
        self._gather = None

    def add_callback(self, event, func):
        self.callbacks[event].append(func)

    def run_callbacks(self, event):
        for cb in self.callbacks.get(event, []):
            cb(self)

    def _dist(self):
        """(world, rank) when batches are sharded over the ranks of an initialised process group, else (1, 0)."""
        import torch.distributed as dist

        if getattr(self.args, "distributed", True) and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist.get_world_size(), dist.get_rank()
        return 1, 0

    def engine_for(self, B, H, W, dtype=torch.float32) -> Engine:
        a = self.args
        key = (B, H, W, dtype, a.conf, a.iou, a.max_det, tuple(a.classes) if a.classes is not None else None, a.agnostic_nms,
               a.multi_label, a.micro_batch, a.cuda_graph)
        eng = self.engines.get(key)
        if eng is not None:
            self.engines.move_to_end(key)
            return eng
        # engines hold their arenas: bounded by count and by bytes; the least recently used one goes first (never one with a
        # batch in flight: `_inflight` is drained before anything is dropped)
        limit = max(int(getattr(a, "engine_cache", 8)), 1)
        while self.engines and (len(self.engines) >= limit or sum(e.plan.arena_bytes for e in self.engines.values()) > (48 << 30)):
            self._flush()
            self.engines.popitem(last=False)
        eng = Engine(self.model, B, (H, W), self.device, micro_batch=a.micro_batch, conf=a.conf, iou=a.iou,
                     max_det=a.max_det, classes=a.classes, agnostic=a.agnostic_nms, multi_label=a.multi_label,
                     cuda_graph=a.cuda_graph, input_dtype=dtype, input_slots=2, rescale=True)
        eng._next_slot = 0
        eng._host = [None, None]                                # pinned (out, counts) per slot, allocated on first use
        eng._slot_done = [None, None]                           # event: the step that last read the slot has finished
        self.engines[key] = eng
        return eng

    def _staging(self, key, shape, device_too=True):
        """Pinned host (+ device) staging pair for raw frames, one per (pipeline slot, position in the batch), grown to the
        largest frame seen and bounded in number (LRU): a folder of mixed-size images no longer leaks a pair per shape."""
        need = int(np.prod(shape))
        ent = self._pinned.get(key)
        if ent is None or ent[0].numel() < need:
            host = torch.empty((need,), dtype=torch.uint8).pin_memory()
            dev = torch.empty((need,), dtype=torch.uint8, device=self.device) if device_too else None
            ent = (host, dev)
            self._pinned[key] = ent
            while len(self._pinned) > 4 * 256:                  # 2 slots x up to 512 positions
                self._pinned.popitem(last=False)
        else:
            self._pinned.move_to_end(key)
        host, dev = ent
        return host[:need].view(shape), (dev[:need].view(shape) if dev is not None else None)

    # ---- source handling (reference data/build.py:186-219, data/loaders.py) ---------------------
    def _batches(self, source):
        """Yield (paths, im0s, tensor_or_None): a Tensor source is ONE batch (loaders.py:575-580); a list / ndarray /
        path source is grouped by args.batch."""
        if isinstance(source, torch.Tensor):
            im = source[None] if source.dim() == 3 else source
            if im.shape[2] % 32 or im.shape[3] % 32:
                raise ValueError(f"tensor source {tuple(im.shape)} must have H, W divisible by stride 32 (loaders.py:559)")
            if im.dtype.is_floating_point and im.numel() and float(im.max()) > 1.0 + 1e-5:
                print("WARNING: torch.Tensor inputs should be normalized 0.0-1.0; dividing by 255 (loaders.py:565-571)")
                im = im.float() / 255.0
            yield [f"image{i}.jpg" for i in range(im.shape[0])], None, im
            return
        if hasattr(source, "__next__"):
            yield from self._iter_batches(source)
            return
        items = source if isinstance(source, (list, tuple)) else [source]
        if items and all(isinstance(s, (str, Path)) for s in items):
            yield from self._file_batches(source)
            return
        imgs, paths = [], []
        for i, s in enumerate(items):
            if isinstance(s, (str, Path)):
                raise TypeError("a list source holds either paths or in-memory images, not both (data/build.py:186-219)")
            if isinstance(s, np.ndarray):
                imgs.append(s)
                paths.append(f"image{i}.jpg")
            else:  # PIL
                imgs.append(np.asarray(s)[:, :, ::-1] if np.asarray(s).ndim == 3 else np.asarray(s))
                paths.append(getattr(s, "filename", "") or f"image{i}.jpg")
        yield paths, imgs, None                                   # in-memory images: ONE batch (LoadPilAndNumpy, loaders.py:451-513)

    def _iter_batches(self, it):
        """A frame stream (any iterator / generator), consumed as it is needed like the reference's stream loaders
        (data/loaders.py:32-220): uint8 HWC BGR frames are grouped `args.batch` at a time (a change of frame shape closes the
        batch early, so that every batch keeps the batched staging path); a 4-D tensor or a list of frames is one batch."""
        from itertools import islice

        bs = max(int(self.args.batch), 1)
        paths, imgs, n = [], [], 0
        it = iter(it)
        while True:
            if not imgs:
                # fast path: a full batch of equally shaped ndarray frames (the common stream) costs one islice and one set of
                # shapes instead of five Python operations per frame (512 frames per step on rank 0 of an 8-GPU run)
                chunk = list(islice(it, bs))
                if not chunk:
                    break
                if len(chunk) == bs and all(type(c) is np.ndarray for c in chunk) and len({c.shape for c in chunk}) == 1:
                    yield [f"image{k}.jpg" for k in range(n, n + bs)], chunk, None
                    n += bs
                    continue
            else:
                chunk = list(islice(it, 1))
                if not chunk:
                    break
            for item in chunk:
                if isinstance(item, torch.Tensor) or isinstance(item, (list, tuple)):
                    if imgs:
                        yield paths, imgs, None
                        paths, imgs = [], []
                    if isinstance(item, torch.Tensor):
                        yield from self._batches(item)
                    else:
                        yield [f"image{n + i}.jpg" for i in range(len(item))], list(item), None
                        n += len(item)
                    continue
                if not isinstance(item, np.ndarray):
                    item = np.asarray(item)[:, :, ::-1] if np.asarray(item).ndim == 3 else np.asarray(item)     # PIL
                if imgs and item.shape != imgs[0].shape:
                    yield paths, imgs, None
                    paths, imgs = [], []
                imgs.append(item)
                paths.append(f"image{n}.jpg")
                n += 1
                if len(imgs) == bs:
                    yield paths, imgs, None
                    paths, imgs = [], []
        if imgs:
            yield paths, imgs, None

    def _file_batches(self, source):
        """Paths -> batches of `args.batch` decoded frames (LoadImagesAndVideos, data/loaders.py:284-446): a `.txt` list, a
        list (sorted), a glob, a directory (`*.*`) or a file; images first, then videos (frames every `vid_stride`, read as they
        are needed); an image batch never runs into the videos; unreadable images are skipped with a warning."""
        import glob
        import os

        import cv2

        parent = None
        if isinstance(source, (str, Path)) and Path(source).suffix == ".txt":
            parent = Path(source).parent
            source = Path(source).read_text().splitlines()
        files = []
        for s in sorted(str(x) for x in source) if isinstance(source, (list, tuple)) else [str(source)]:
            a = str(Path(s).absolute())
            if "*" in a:
                files.extend(sorted(glob.glob(a, recursive=True)))
            elif os.path.isdir(a):
                files.extend(sorted(glob.glob(os.path.join(a, "*.*"))))
            elif os.path.isfile(a):
                files.append(a)
            elif parent is not None and (parent / s).is_file():
                files.append(str((parent / s).absolute()))
            else:
                raise FileNotFoundError(f"{s} does not exist")
        ext = lambda f: f.rsplit(".", 1)[-1].lower()              # noqa: E731
        images = [f for f in files if ext(f) in IMG_FORMATS]
        videos = [f for f in files if ext(f) in VID_FORMATS]
        if not images and not videos:
            raise FileNotFoundError(f"No images or videos found in {source}")
        bs = max(int(self.args.batch), 1)
        stride = max(int(getattr(self.args, "vid_stride", 1)), 1)
        paths, imgs = [], []
        for f in images:
            im = cv2.imread(f)
            if im is None:
                print(f"WARNING: image read error {f}")
                continue
            paths.append(f)
            imgs.append(im)
            if len(imgs) == bs:
                yield paths, imgs, None
                paths, imgs = [], []
        if imgs:
            yield paths, imgs, None
            paths, imgs = [], []
        for f in videos:
            cap = cv2.VideoCapture(f)
            if not cap.isOpened():
                raise FileNotFoundError(f"Failed to open video {f}")
            try:
                while True:
                    ok = False
                    for _ in range(stride):
                        ok = cap.grab()
                        if not ok:
                            break
                    if not ok:
                        break
                    ok, im = cap.retrieve()
                    if not ok:
                        break
                    paths.append(f)
                    imgs.append(im)
                    if len(imgs) == bs:
                        yield paths, imgs, None
                        paths, imgs = [], []
            finally:
                cap.release()
        if imgs:
            yield paths, imgs, None

    # ---- preprocess (predictor.py:118-163) -----------------------------------------------------------
    def preprocess(self, im0s):
        """uint8 HWC BGR list -> uint8 (B,3,H,W) RGB batch on the host (predictor.py:118-131, 147-163): the reference's own
        host path (cv2 LetterBox + stack + BGR->RGB + HWC->CHW); its `.float()` and `/ 255` (:133-135) happen in the stem
        kernel.  Used when `gpu_preprocess=False` or for frames the GPU letterbox does not take."""
        a = self.args
        shape = (a.imgsz, a.imgsz) if isinstance(a.imgsz, int) else tuple(a.imgsz)
        same = len({x.shape for x in im0s}) == 1
        lb = [letterbox(x, shape, auto=same) for x in im0s]
        return torch.from_numpy(np.ascontiguousarray(np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2)))

    def _canvas(self, im0s):
        """(H, W, per-frame geometry or None) of the letterboxed batch; geometry None -> host path."""
        a = self.args
        shape = (a.imgsz, a.imgsz) if isinstance(a.imgsz, int) else tuple(a.imgsz)
        shapes = {x.shape for x in im0s}
        same = len(shapes) == 1
        if same:                                                           # one geometry for the whole batch (512 frames per step with 8 ranks)
            geo = [letterbox_geometry(im0s[0].shape[:2], shape, auto=True)] * len(im0s)
        else:
            geo = [letterbox_geometry(x.shape[:2], shape, auto=False) for x in im0s]
        H, W = geo[0][4], geo[0][5]
        ok = (getattr(a, "gpu_preprocess", True) and W % 4 == 0 and all((g[4], g[5]) == (H, W) for g in geo)
              and all(x.ndim == 3 and x.shape[2] == 3 and x.dtype == np.uint8 for x in im0s))
        return H, W, (geo if ok else None)

    def _upload_raw(self, eng, slot, im0s, geo, lo, hi):
        """Frames lo..hi -> device staging -> letterbox kernel -> input slot, all on the copy stream.  A frame that already
        lives in pinned host memory (e.g. a view of a capture ring) is uploaded from where it is and must stay unchanged
        until its Results arrive; any other frame goes through a pinned staging buffer, filled by a few host threads
        (a single-threaded memcpy of a 64 x 640 x 640 batch costs more than the whole device step)."""
        n = hi - lo
        if n <= 0:
            return
        cs = self._copy_stream
        frames = [im0s[j] if im0s[j].flags.c_contiguous else np.ascontiguousarray(im0s[j]) for j in range(lo, hi)]
        same = all(f.shape == frames[0].shape for f in frames) and all(geo[j] == geo[lo] for j in range(lo, hi))
        pinned = [torch.from_numpy(f).is_pinned() if f.flags.writeable else False for f in frames]
        if same:
            host, dev = self._staging(("rawb", slot), (n,) + frames[0].shape)
            hosts, devs = list(host), list(dev)
        else:
            pairs = [self._staging(("raw", slot, k), f.shape) for k, f in enumerate(frames)]
            hosts, devs = [h for h, _ in pairs], [d for _, d in pairs]
        todo = [k for k in range(n) if not pinned[k]]
        if len(todo) > 1:
            list(self._pool().map(lambda k: np.copyto(hosts[k].numpy(), frames[k]), todo))
        elif todo:
            np.copyto(hosts[todo[0]].numpy(), frames[todo[0]])
        with torch.cuda.stream(cs):
            if same and not any(pinned):
                dev.copy_(host, non_blocking=True)
            else:
                for k in range(n):
                    devs[k].copy_(torch.from_numpy(frames[k]) if pinned[k] else hosts[k], non_blocking=True)
            if same:
                g = geo[lo]
                K.letterbox_u8_batch(dev, eng.image_slots[slot][:n], g[0], g[1], g[2], g[3])
            else:
                for k in range(n):
                    g = geo[lo + k]
                    K.letterbox_u8(devs[k], eng.image_slots[slot][k], g[0], g[1], g[2], g[3])

    def _pool(self):
        if getattr(self, "_threads", None) is None:
            import os
            from concurrent.futures import ThreadPoolExecutor

            self._threads = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) - 1)))
        return self._threads

    # ---- the loop (predictor.py:221-306) ------------------------------------------------------------------
    def __call__(self, source=None, model=None, stream=False):
        gen = self.stream_inference(source, model)
        return gen if stream else list(gen)

    def _launch(self, paths, im0s, tensor):
        """Stage + upload one batch into a free input slot (copy stream), enqueue its engine step and the D2H of its padded
        detections (compute stream).  Nothing here waits for the device."""
        world, rank = self._dist()
        nvtx = torch.cuda.nvtx
        nvtx.range_push("dy.preprocess")          # staging + upload + letterbox (copy stream)
        n_items = tensor.shape[0] if tensor is not None else len(im0s)
        lo, hi = (0, n_items) if world == 1 else shard_bounds(n_items, world, rank)
        b_local = n_items if world == 1 else -(-n_items // world)            # equal shards (the last ones padded): one collective
        t0 = time.perf_counter()
        compute = torch.cuda.current_stream(self.device)
        cs = self._copy_stream
        if tensor is not None:
            im = tensor if tensor.dtype == torch.uint8 else tensor.float()
            H, W = im.shape[2:]
            eng = self.engine_for(b_local, H, W, im.dtype)
            slot = eng._next_slot
            if eng._slot_done[slot] is not None:
                cs.wait_event(eng._slot_done[slot])
            if im.is_cuda:
                cs.wait_stream(torch.cuda.current_stream(im.device))       # the caller's tensor may still be being written
            with torch.cuda.stream(cs):
                if hi > lo:
                    eng.image_slots[slot][: hi - lo].copy_(im[lo:hi], non_blocking=True)
            oshapes = [(H, W)] * n_items
            in_shape = (H, W)
        else:
            H, W, geo = self._canvas(im0s)
            eng = self.engine_for(b_local, H, W, torch.uint8)
            slot = eng._next_slot
            if eng._slot_done[slot] is not None:
                cs.wait_event(eng._slot_done[slot])
            if geo is None:                                                # host letterbox (the reference's own path)
                arr = self.preprocess(im0s[lo:hi]) if hi > lo else None
                if arr is not None and tuple(arr.shape[2:]) != (H, W):
                    raise _C.DroneYoloError("host letterbox produced another canvas than the planned one")
                with torch.cuda.stream(cs):
                    if arr is not None:
                        host, _ = self._staging(("canvas", slot), tuple(arr.shape), device_too=False)
                        host.copy_(arr)
                        eng.image_slots[slot][: hi - lo].copy_(host, non_blocking=True)
            else:
                # raw frames cross PCIe as they are; the letterbox kernel does resize + border + BGR->RGB + HWC->CHW straight
                # into the engine's input slot (dy_letterbox_u8): ONE launch for a batch of equally shaped frames
                self._upload_raw(eng, slot, im0s, geo, lo, hi)
            oshapes = [tuple(im0s[0].shape[:2])] * n_items if geo is not None and geo[0] is geo[-1] else [tuple(o.shape[:2]) for o in im0s]
            in_shape = (H, W)
        # per-image scale_boxes / clip_boxes parameters for the NMS output phase (this rank's shard)
        rs = ops.rescale_params(in_shape, oshapes[lo:hi] + [in_shape] * (b_local - (hi - lo)))
        rs_host, _ = self._staging(("rescale", slot, b_local), (b_local * 8 * 4,), device_too=False)
        rs_host.view(torch.float32).view(b_local, 8).copy_(rs)
        with torch.cuda.stream(cs):
            eng.rescale_slots[slot].copy_(rs_host.view(torch.float32).view(b_local, 8), non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(cs)
        eng._next_slot = slot ^ 1
        t1 = time.perf_counter()
        compute.wait_event(uploaded)
        nvtx.range_pop()
        nvtx.range_push("dy.inference")           # conv stack + decode + NMS (+ rescale): one graph replay
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(compute)
        out, counts = eng.step(slot)
        e1.record(compute)
        nvtx.range_pop()
        if world > 1:
            if self._gather is None or self._gather.key != (b_local, self.args.max_det):
                self._gather = DetectionGather(b_local, self.args.max_det, self.device)
            out, counts = self._gather.gather(out, counts)
        host_pair = eng._host[slot]
        if host_pair is None or host_pair[0].shape != out.shape:
            host_pair = eng._host[slot] = (torch.empty(out.shape, dtype=torch.float32).pin_memory(),
                                           torch.empty(counts.shape, dtype=torch.int32).pin_memory())
        if rank == 0:
            host_pair[0].copy_(out, non_blocking=True)
            host_pair[1].copy_(counts, non_blocking=True)
        done = torch.cuda.Event()
        done.record(compute)
        eng._slot_done[slot] = done                                        # the upload after next into this slot waits for it
        return dict(paths=paths, im0s=im0s, tensor=tensor, eng=eng, slot=slot, host=host_pair, done=done, e0=e0, e1=e1,
                    pre_ms=(t1 - t0) * 1e3, n=n_items, in_shape=in_shape, oshapes=oshapes, world=world, rank=rank, b_local=b_local)

    def _finish(self, rec):
        """Wait for ONE batch (its own event, not the device), build Results on the host, run the callbacks."""
        rec["done"].synchronize()
        t2 = time.perf_counter()
        torch.cuda.nvtx.range_push("dy.postprocess")  # Results on the host
        self.batch = (rec["paths"], rec["im0s"], None)
        if rec["rank"] != 0:
            self.results = []                                              # rank 0 holds every detection of the batch
        else:
            self.results = self.construct_results(rec)
        t3 = time.perf_counter()
        torch.cuda.nvtx.range_pop()
        self.run_callbacks("on_predict_postprocess_end")
        n = max(rec["n"], 1)
        speed = {"preprocess": rec["pre_ms"] / n, "inference": rec["e0"].elapsed_time(rec["e1"]) / n, "postprocess": (t3 - t2) * 1e3 / n}
        for r in self.results:
            r.speed = speed
        self.seen += rec["n"]
        if self.args.verbose:
            for p, r in zip(rec["paths"], self.results):
                print(f"{p}: {rec['in_shape'][0]}x{rec['in_shape'][1]} {r.verbose()}{speed['inference']:.2f}ms")
        self.run_callbacks("on_predict_batch_end")
        return self.results

    def _flush(self):
        """Drop the batches in flight (an engine is about to be evicted or the model replaced)."""
        q = getattr(self, "_inflight", None)
        if q:
            for rec in q:
                rec["done"].synchronize()
            q.clear()

    def stream_inference(self, source=None, model=None):
        if self.model is None:
            self.setup_model(model)
        self._check_args()
        with self._lock, torch.inference_mode(), torch.cuda.device(self.device):
            self.run_callbacks("on_predict_start")
            self._inflight = deque()
            for paths, im0s, tensor in self._batches(source):
                self.run_callbacks("on_predict_batch_start")
                self.batch = (paths, im0s, None)
                self._inflight.append(self._launch(paths, im0s, tensor))
                if len(self._inflight) == 2:                               # batch i+1 is queued: turn batch i into Results meanwhile
                    yield from self._finish(self._inflight.popleft())
            while self._inflight:
                yield from self._finish(self._inflight.popleft())
            self.run_callbacks("on_predict_end")

    def inference(self, eng: Engine, slot: int = 0):
        """conv stack + decode + NMS, one CUDA-graph replay (predictor.py:138-145 + detect/predict.py:25-35)."""
        return eng.step(slot)

    def construct_results(self, rec):
        """Padded detections (already in original-image coordinates: the NMS output phase applied scale_boxes + clip_boxes)
        -> Results (detect/predict.py:37-73).  With 8 ranks, rank 0 builds 512 Results per engine step: everything per image
        is kept to a few hundred nanoseconds - one copy of the block in image order, and every Results slices its own rows out
        of it when its boxes are first touched (a Python-level slice per image costs 2 us, the shard lookup 3 us)."""
        host, cnt = rec["host"]
        world, b_local, n_items = rec["world"], rec["b_local"], rec["n"]
        data, counts = host.numpy(), cnt.numpy()
        md = data.shape[1]
        key = (n_items, world, b_local, md)
        if getattr(self, "_perm_key", None) != key:                        # image i lives at row perm[i] of the (gathered) block
            perm = np.arange(n_items, dtype=np.int64)
            for r in range(world if world > 1 else 0):
                lo, hi = shard_bounds(n_items, world, r)
                perm[lo:hi] = r * b_local + np.arange(hi - lo)
            self._perm_key, self._perm = key, perm
        # ONE copy of the (gathered) block in image order - the pinned pair is reused two batches later - shared by the batch's
        # Results; each takes its (image, row count) view when its boxes are first touched.  numpy, not torch: CPU torch ops of this
        # size fan out to the intra-op thread pool and take milliseconds
        block = torch.from_numpy(np.take(data, self._perm, axis=0) if world > 1 else data[:n_items].copy())
        rows = list(zip([block] * n_items, range(n_items), counts[self._perm].tolist()))
        names = self.model.names
        orig_imgs, tensor, paths, oshapes = rec["im0s"], rec["tensor"], rec["paths"], rec["oshapes"]
        if tensor is not None:
            def lazy(j, t=tensor):
                def orig():
                    if t.dtype == torch.uint8:
                        return t[j].permute(1, 2, 0).contiguous().cpu().numpy()
                    return ops.convert_torch2numpy_batch(t[j:j + 1].float())[0]
                return orig
            return [Results(lazy(i), path=paths[i], names=names, boxes=rows[i], orig_shape=oshapes[i]) for i in range(n_items)]
        return [Results(o, path=pt, names=names, boxes=rw, orig_shape=sh) for o, pt, rw, sh in zip(orig_imgs, paths, rows, oshapes)]

    def postprocess(self, preds, img, orig_imgs, paths):
        """The reference's hook (detect/predict.py:23-73) for padded detections in INPUT-pixel coordinates (an engine built
        without the fused post-step, or `ops.nms_padded`): one D2H of (B, max_det, 6), one batched scale_boxes + clip_boxes
        (the per-image arithmetic, element for element), Results."""
        out, counts = preds
        B = out.shape[0]
        host = out.cpu() if out.is_cuda else out.clone()
        n = counts.cpu().tolist()
        names = self.model.names
        in_shape = tuple(img.shape[2:])
        tensor_src = isinstance(orig_imgs, torch.Tensor)
        oshapes = [in_shape] * B if tensor_src else [tuple(o.shape[:2]) for o in orig_imgs]
        ops.scale_boxes_batch(in_shape, host[..., :4], oshapes)
        results = []
        for i in range(B):
            if tensor_src:
                def orig(j=i, t=orig_imgs):
                    return ops.convert_torch2numpy_batch(t[j:j + 1])[0]
            else:
                orig = orig_imgs[i]
            results.append(Results(orig, path=paths[i], names=names, boxes=host[i, :n[i]], orig_shape=oshapes[i]))
        return results

    def predict_cli(self, source=None, model=None):
        for _ in self.stream_inference(source, model):
            pass
