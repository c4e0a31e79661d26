"""DetectionPredictor: source -> preprocess -> Engine (conv stack + decode + NMS) -> Results.

Mirror of the reference's BasePredictor / DetectionPredictor (ultralytics/engine/predictor.py:66-410,
models/yolo/detect/predict.py:23-73) for the inference hot path: same constructor (`overrides`, `_callbacks`),
`setup_model`, `__call__(source, stream)`, `preprocess`, `inference`, `postprocess`, the five `on_predict_*`
callback events and `Results.speed`.  Differences: one batched D2H of the padded detections instead of per-image
boolean-mask syncs; `orig_img` of tensor sources is produced lazily; boxes are rescaled on the host (<= max_det rows).
"""
from __future__ import annotations

import threading
import time
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from .. import _C
from .. import kernels as K
from ..utils import ops
from .engine import Engine
from .results import Results

DEFAULTS = dict(task="detect", mode="predict", imgsz=640, batch=1, device=None, conf=0.25, iou=0.7, max_det=300,
                half=False, classes=None, agnostic_nms=False, augment=False, stream=False, verbose=False,
                micro_batch=0, cuda_graph=True, multi_label=False, gpu_preprocess=True, engine_cache=8, vid_stride=1)
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
        if args.get("augment"):
            raise _C.DroneYoloError("augment=True (TTA) is outside the inference hot path")
        self.args = SimpleNamespace(**args)
        self.model = None
        self.device = None
        self.engines: dict = {}
        self.callbacks = {e: [] for e in EVENTS}
        for k, v in (_callbacks or {}).items():
            self.callbacks.setdefault(k, []).extend(v if isinstance(v, (list, tuple)) else [v])
        self.results = None
        self.batch = None
        self.save_dir = None
        self.seen = 0
        self._lock = threading.Lock()
        self._pinned: dict = {}

    # ---- setup ---------------------------------------------------------------------------------
    def setup_model(self, model, verbose=False):
        """Move to the device, fold BN + re-parameterise (AutoBackend: autobackend.py:149-159)."""
        dev = self.args.device
        if dev is None or dev == "":
            dev = "cuda:0"
        if isinstance(dev, int) or (isinstance(dev, str) and dev.isdigit()):
            dev = f"cuda:{dev}"
        self.device = torch.device(dev)
        if self.device.type != "cuda":
            raise _C.DroneYoloError(f"device '{dev}': drone_yolo_b200 has no CPU path (use the reference for CPU inference)")
        self.model = model.to(self.device).eval().fuse(verbose=False)
        self.engines.clear()

    def add_callback(self, event, func):
        self.callbacks[event].append(func)

    def run_callbacks(self, event):
        for cb in self.callbacks.get(event, []):
            cb(self)

    def engine_for(self, B, H, W, dtype=torch.float32) -> Engine:
        a = self.args
        key = (B, H, W, dtype, a.conf, a.iou, a.max_det, tuple(a.classes) if a.classes else None, a.agnostic_nms, a.multi_label,
               a.micro_batch, a.cuda_graph)
        if key not in self.engines:
            held = sum(e.plan.arena_bytes for e in self.engines.values())            # engines hold their arenas: bounded by
            if len(self.engines) >= max(int(getattr(a, "engine_cache", 8)), 1) or held > (48 << 30):   # count and by bytes
                self.engines.clear()
            self.engines[key] = Engine(self.model, B, (H, W), self.device, micro_batch=a.micro_batch, conf=a.conf, iou=a.iou,
                                       max_det=a.max_det, classes=a.classes, agnostic=a.agnostic_nms,
                                       multi_label=a.multi_label, cuda_graph=a.cuda_graph, input_dtype=dtype)
        return self.engines[key]

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

    def preprocess(self, im0s):
        """uint8 HWC BGR list -> pinned uint8 (B,3,H,W) RGB (predictor.py:118-131, 147-163).  Like the reference, the
        batch crosses PCIe as uint8; its `.float()` and `/ 255` (predictor.py:133-135) happen on the device, here inside the
        stem kernel."""
        a = self.args
        shape = (a.imgsz, a.imgsz) if isinstance(a.imgsz, int) else tuple(a.imgsz)
        same = len({x.shape for x in im0s}) == 1
        if getattr(a, "gpu_preprocess", True) and all(x.ndim == 3 and x.shape[2] == 3 and x.dtype == np.uint8 for x in im0s):
            return self.preprocess_gpu(im0s, shape, same)
        lb = [letterbox(x, shape, auto=same) for x in im0s]
        arr = np.ascontiguousarray(np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2))
        key = arr.shape
        if key not in self._pinned:
            self._pinned[key] = torch.empty(key, dtype=torch.uint8).pin_memory()
        buf = self._pinned[key]
        buf.copy_(torch.from_numpy(arr))
        return buf

    def preprocess_gpu(self, im0s, shape, same):
        """The raw frames cross PCIe as they are (pinned staging, one buffer per frame size) and ONE kernel per frame does
        resize + border + BGR->RGB + HWC->CHW straight into the engine's uint8 input batch (dy_letterbox_u8): the host's
        cv2.resize / copyMakeBorder / stack / transpose (about 1 ms per 1080p frame on one core) disappear."""
        geo = [letterbox_geometry(x.shape[:2], shape, auto=same) for x in im0s]
        H, W = geo[0][4], geo[0][5]
        if any((g[4], g[5]) != (H, W) for g in geo) or W % 4:
            lb = [letterbox(x, shape, auto=same) for x in im0s]          # ragged canvases: the host path decides
            return torch.from_numpy(np.ascontiguousarray(np.stack(lb)[..., ::-1].transpose(0, 3, 1, 2)))
        eng = self.engine_for(len(im0s), H, W, torch.uint8)
        for i, (x, g) in enumerate(zip(im0s, geo)):
            x = np.ascontiguousarray(x)
            key = ("raw", i, x.shape)
            if key not in self._pinned:
                self._pinned[key] = (torch.empty(x.shape, dtype=torch.uint8).pin_memory(),
                                     torch.empty(x.shape, dtype=torch.uint8, device=self.device))
            host, dev = self._pinned[key]
            host.copy_(torch.from_numpy(x))
            dev.copy_(host, non_blocking=True)
            K.letterbox_u8(dev, eng.images[i], g[0], g[1], g[2], g[3])
        return eng.images

    # ---- the loop --------------------------------------------------------------------------------
    def __call__(self, source=None, model=None, stream=False):
        gen = self.stream_inference(source, model)
        return gen if stream else list(gen)

    def stream_inference(self, source=None, model=None):
        if self.model is None:
            self.setup_model(model)
        with self._lock, torch.inference_mode():
            self.run_callbacks("on_predict_start")
            for paths, im0s, tensor in self._batches(source):
                self.run_callbacks("on_predict_batch_start")
                self.batch = (paths, im0s, None)
                t0 = time.perf_counter()
                im = tensor if tensor is not None else self.preprocess(im0s)
                B, _, H, W = im.shape
                if im.dtype != torch.uint8:
                    im = im.float()
                eng = self.engine_for(B, H, W, im.dtype)
                if im.data_ptr() != eng.images.data_ptr():         # the GPU preprocess writes the static input in place
                    eng.images.copy_(im, non_blocking=True)        # H2D (or D2D) into the static input
                torch.cuda.synchronize(self.device)
                t1 = time.perf_counter()
                out, counts = self.inference(eng)
                torch.cuda.synchronize(self.device)
                t2 = time.perf_counter()
                self.results = self.postprocess((out, counts), im, im0s if im0s is not None else tensor, paths)
                t3 = time.perf_counter()
                self.run_callbacks("on_predict_postprocess_end")
                n = len(self.results)
                speed = {"preprocess": (t1 - t0) * 1e3 / n, "inference": (t2 - t1) * 1e3 / n, "postprocess": (t3 - t2) * 1e3 / n}
                for r in self.results:
                    r.speed = speed
                self.seen += n
                if self.args.verbose:
                    for p, r in zip(paths, self.results):
                        print(f"{p}: {H}x{W} {r.verbose()}{speed['inference']:.2f}ms")
                self.run_callbacks("on_predict_batch_end")
                yield from self.results
            self.run_callbacks("on_predict_end")

    def inference(self, eng: Engine):
        """conv stack + decode + NMS, one CUDA-graph replay (predictor.py:138-145 + detect/predict.py:25-35)."""
        return eng.step()

    def postprocess(self, preds, img, orig_imgs, paths):
        """Padded detections -> Results; boxes rescaled to the original image (detect/predict.py:37-73).  One D2H of
        (B, max_det, 6), one batched rescale + clamp (the reference's per-image scale_boxes arithmetic, element for element);
        each Results holds a view of its first counts[i] rows."""
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
