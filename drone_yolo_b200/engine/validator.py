"""DetectionValidator: the validator-side caller of the hot path.

Reference: `ultralytics/engine/validator.py:101-222` (`BaseValidator.__call__`: preprocess -> model -> postprocess ->
update_metrics per batch, then get_stats) and `models/yolo/detect/val.py:30-182` (`DetectionValidator`: preprocess with the
`save_hybrid` a-priori labels, `postprocess` = NMS at conf 0.001 with `multi_label=True` and `labels=self.lb`, native-space
rescale of predictions and labels, the (N, 10) correct matrix, AP per class).

What runs on the GPU here: the conv stack + decode (`Engine.enqueue(nms=False)` on the uint8 / float batch, resident in an
engine input slot) and the multi-label NMS (`ops.non_max_suppression` -> `dy_nms`, bit-exact against the reference's rows).
What stays on the host, as in the reference: matching the <= max_det rows per image against the labels and the AP
arithmetic (`utils/metrics.py`).  Datasets / dataloaders are outside the hot path: `__call__` takes any iterable of batch
dicts shaped like the reference dataloader's (`img` uint8 (B,3,H,W), `cls` (n,1), `bboxes` (n,4) normalised xywh,
`batch_idx` (n,), `ori_shape`, `ratio_pad`).
"""
from __future__ import annotations

import time
from types import SimpleNamespace

import numpy as np
import torch

from .._C import DroneYoloError
from ..utils import ops
from ..utils.metrics import DetMetrics, box_iou, match_predictions
from .engine import Engine

_DEFAULTS = dict(conf=None, iou=0.7, max_det=300, half=False, save_hybrid=False, single_cls=False, agnostic_nms=False,
                 imgsz=640, batch=16, device=None, task="detect", plots=False, save_json=False, save_txt=False, micro_batch=0,
                 cuda_graph=True)


class DetectionValidator:
    def __init__(self, dataloader=None, save_dir=None, pbar=None, args=None, _callbacks=None):
        a = dict(_DEFAULTS)
        a.update(args or {})
        self.args = SimpleNamespace(**a)
        if self.args.conf is None:
            self.args.conf = 0.001                                   # validator.py:101-102
        if self.args.half:
            raise DroneYoloError("half=True: the conv stack always computes in bf16; there is no fp16 / fp32 switch")
        if self.args.plots or self.args.save_json or self.args.save_txt:
            raise DroneYoloError("plots / save_json / save_txt are outside the inference hot path")
        self.dataloader = dataloader
        self.iouv = torch.linspace(0.5, 0.95, 10)                    # val.py:41
        self.niou = self.iouv.numel()
        self.lb = []                                                 # a-priori labels for autolabelling (val.py:43)
        self.metrics = DetMetrics()
        self.speed = {"preprocess": 0.0, "inference": 0.0, "loss": 0.0, "postprocess": 0.0}
        self.device = None
        self.model = None
        self._engines = {}
        self.seen = 0
        self.stats = None

    # ------------------------------------------------------------------------------------------------------------
    def init_metrics(self, model):
        """val.py:68-87 without the dataset-name branches."""
        self.names = getattr(model, "names", None) or {i: str(i) for i in range(model.model[-1].nc)}
        self.nc = len(self.names)
        self.metrics.names = self.names
        self.seen = 0
        self.stats = dict(tp=[], conf=[], pred_cls=[], target_cls=[], target_img=[])

    def preprocess(self, batch):
        """val.py:50-66: images to the device (kept uint8: the stem kernel applies 1/255 itself, exactly), labels to the
        device, and with `save_hybrid` the per-image (cls, x, y, w, h) pixel label rows NMS will append."""
        img = batch["img"]
        if not isinstance(img, torch.Tensor):
            img = torch.from_numpy(np.ascontiguousarray(img))
        if img.dtype not in (torch.uint8, torch.float32):
            raise DroneYoloError(f"validator images must be uint8 0..255 (or float32 already in [0,1]), got {img.dtype}")
        batch["img"] = img.to(self.device, non_blocking=True)
        for k in ("batch_idx", "cls", "bboxes"):
            batch[k] = torch.as_tensor(batch[k]).to(self.device)
        if self.args.save_hybrid:
            height, width = batch["img"].shape[2:]
            nb = len(batch["img"])
            bboxes = batch["bboxes"] * torch.tensor((width, height, width, height), device=self.device)
            self.lb = [torch.cat([batch["cls"][batch["batch_idx"] == i], bboxes[batch["batch_idx"] == i]], dim=-1) for i in range(nb)]
        return batch

    def _engine(self, B, H, W, dtype):
        key = (B, H, W, dtype)
        eng = self._engines.get(key)
        if eng is None:
            if len(self._engines) >= 4:
                self._engines.pop(next(iter(self._engines)))
            eng = self._engines[key] = Engine(self.model, B, (H, W), self.device, micro_batch=self.args.micro_batch, conf=self.args.conf,
                                              iou=self.args.iou, max_det=self.args.max_det, multi_label=True,
                                              agnostic=self.args.single_cls or self.args.agnostic_nms, cuda_graph=False,
                                              input_dtype=dtype)
        return eng

    def inference(self, img):
        """The model call of validator.py:180: conv stack + decode on the GPU -> the (B, 4+nc, A) prediction tensor."""
        B, _, H, W = img.shape
        eng = self._engine(B, H, W, img.dtype)
        eng.images.copy_(img, non_blocking=True)
        eng.enqueue(nms=False)
        return eng.y

    def postprocess(self, preds):
        """val.py:93-106."""
        return ops.non_max_suppression(preds, self.args.conf, self.args.iou, labels=self.lb, nc=self.nc, multi_label=True,
                                       agnostic=self.args.single_cls or self.args.agnostic_nms, max_det=self.args.max_det)

    # ------------------------------------------------------------------------------------------------------------
    def _prepare_batch(self, si, batch):
        """val.py:108-119: this image's labels in native (original-image) pixels."""
        idx = batch["batch_idx"] == si
        cls = batch["cls"][idx].squeeze(-1)
        bbox = batch["bboxes"][idx]
        ori_shape = batch["ori_shape"][si]
        imgsz = batch["img"].shape[2:]
        ratio_pad = batch["ratio_pad"][si]
        if len(cls):
            bbox = ops.xywh2xyxy(bbox) * torch.tensor(imgsz, device=self.device)[[1, 0, 1, 0]]
            ops.scale_boxes(imgsz, bbox, ori_shape, ratio_pad=ratio_pad)
        return {"cls": cls, "bbox": bbox, "ori_shape": ori_shape, "imgsz": imgsz, "ratio_pad": ratio_pad}

    def _prepare_pred(self, pred, pbatch):
        """val.py:121-127."""
        predn = pred.clone()
        ops.scale_boxes(pbatch["imgsz"], predn[:, :4], pbatch["ori_shape"], ratio_pad=pbatch["ratio_pad"])
        return predn

    def _process_batch(self, detections, gt_bboxes, gt_cls):
        """val.py:213-232."""
        return match_predictions(detections[:, 5], gt_cls, box_iou(gt_bboxes, detections[:, :4]), self.iouv)

    def update_metrics(self, preds, batch):
        """val.py:129-175 (no confusion matrix / json / txt)."""
        for si, pred in enumerate(preds):
            self.seen += 1
            npr = len(pred)
            stat = dict(conf=torch.zeros(0, device=self.device), pred_cls=torch.zeros(0, device=self.device),
                        tp=torch.zeros(npr, self.niou, dtype=torch.bool, device=self.device))
            pbatch = self._prepare_batch(si, batch)
            cls, bbox = pbatch.pop("cls"), pbatch.pop("bbox")
            nl = len(cls)
            stat["target_cls"] = cls
            stat["target_img"] = cls.unique()
            if npr == 0:
                if nl:
                    for k in self.stats:
                        self.stats[k].append(stat[k])
                continue
            if self.args.single_cls:
                pred[:, 5] = 0
            predn = self._prepare_pred(pred, pbatch)
            stat["conf"] = predn[:, 4]
            stat["pred_cls"] = predn[:, 5]
            if nl:
                stat["tp"] = self._process_batch(predn, bbox, cls)
            for k in self.stats:
                self.stats[k].append(stat[k])

    def get_stats(self):
        """val.py:177-186."""
        stats = {k: (torch.cat(v, 0).cpu().numpy() if len(v) else np.zeros((0, self.niou) if k == "tp" else (0,))) for k, v in self.stats.items()}
        self.nt_per_class = np.bincount(stats["target_cls"].astype(int), minlength=self.nc)
        self.nt_per_image = np.bincount(stats["target_img"].astype(int), minlength=self.nc)
        stats.pop("target_img", None)
        self.last_stats = stats
        if len(stats) and len(stats["tp"]):
            self.metrics.process(**stats)
        return self.metrics.results_dict

    # ------------------------------------------------------------------------------------------------------------
    def __call__(self, trainer=None, model=None, batches=None):
        """validator.py:101-222 for the not-training case: one pass over `batches` (or the dataloader given at construction)."""
        if trainer is not None:
            raise DroneYoloError("validation inside a training loop is outside the inference hot path")
        net = getattr(model, "model", None)
        net = model if hasattr(model, "fuse") and hasattr(model, "yaml") else net
        if net is None:
            raise DroneYoloError("DetectionValidator needs a DetectionModel (or a YOLO wrapper)")
        dev = torch.device(self.args.device if self.args.device is not None else "cuda:0")
        if dev.type != "cuda":
            raise DroneYoloError("drone_yolo_b200 runs on CUDA (sm_100a) devices only; there is no CPU path")
        self.device = dev
        self.model = net.to(dev).eval().fuse(verbose=False)         # AutoBackend folds BN before validating (autobackend.py:149-159)
        self.init_metrics(self.model)
        data = batches if batches is not None else self.dataloader
        if data is None:
            raise DroneYoloError("DetectionValidator: no batches given (datasets / dataloaders are outside the hot path: pass an iterable of batch dicts)")
        dt = [0.0, 0.0, 0.0]
        with torch.inference_mode(), torch.cuda.device(dev):
            for batch in data:
                t0 = time.perf_counter()
                batch = self.preprocess(dict(batch))
                torch.cuda.synchronize(dev)
                t1 = time.perf_counter()
                preds = self.inference(batch["img"])
                torch.cuda.synchronize(dev)
                t2 = time.perf_counter()
                preds = self.postprocess(preds)
                t3 = time.perf_counter()
                self.update_metrics(preds, batch)
                dt = [dt[0] + t1 - t0, dt[1] + t2 - t1, dt[2] + t3 - t2]
        stats = self.get_stats()
        n = max(self.seen, 1)
        self.speed = dict(zip(("preprocess", "inference", "loss", "postprocess"), (dt[0] / n * 1e3, dt[1] / n * 1e3, 0.0, dt[2] / n * 1e3)))
        self.metrics.speed = self.speed
        return stats
