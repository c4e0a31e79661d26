"""User API: `YOLO(model).predict(source, ...)` (reference ultralytics/engine/model.py:29-560 and
models/yolo/model.py:11-59, detect row of the task map)."""
from __future__ import annotations

import pickle
from pathlib import Path

import torch
import torch.nn as nn

from ..nn.tasks import DetectionModel, attempt_load_state, guess_model_task
from .predictor import DetectionPredictor


class Model(nn.Module):
    def __init__(self, model="yolov8n-p2-repvgg.yaml", task=None, verbose=False, nc=None):
        super().__init__()
        self.predictor = None
        self.model = None
        self.overrides = {}
        self.task = task or "detect"
        self.model_name = str(model)
        if isinstance(model, nn.Module):
            self.model = model
        elif Path(str(model)).suffix in (".yaml", ".yml"):
            self._new(str(model), verbose=verbose, nc=nc)
        else:
            self._load(str(model), nc=nc)

    def _new(self, cfg, verbose=False, nc=None):
        """Build a randomly initialised model from a YAML (engine/model.py:231-264)."""
        self.model = DetectionModel(cfg, nc=nc, verbose=verbose)
        self.task = guess_model_task(self.model)
        self.overrides = {"model": cfg, "task": self.task}

    def _load(self, weights, nc=None):
        """Load `<name>.pt` holding {'yaml': cfg-name-or-dict, 'model': state_dict} or a bare state_dict next to a
        YAML named by `cfg` (engine/model.py:266-302; pickled reference checkpoints: nn/ckpt.py)."""
        try:
            obj = torch.load(weights, map_location="cpu", weights_only=True)
        except pickle.UnpicklingError:
            obj = None                                # not a plain tensor container: a pickled reference checkpoint
        if not (isinstance(obj, dict) and obj.get("yaml") is not None and isinstance(obj.get("model"), dict)):
            # the reference's own format: {'model': <pickled DetectionModel>, 'ema': ..., 'train_args': ...}
            # (nn/tasks.py:786-926, attempt_load_one_weight), read without the reference package (nn/ckpt.py)
            from ..nn.ckpt import load_reference_checkpoint

            cfg, state, names, _ = load_reference_checkpoint(weights)
            obj = {"yaml": cfg, "model": state, "nc": cfg.get("nc", nc)}
            if names is not None:
                obj["names"] = names
        cfg = obj["yaml"]
        self.model = DetectionModel(cfg, nc=obj.get("nc", nc), verbose=False)
        self.model.load(obj["model"], verbose=False)
        got, want, missing = self.model.transferred
        if got != want:
            # the reference logs "Transferred x/y items" and carries on with random weights for the rest (tasks.py:277-279); a
            # predictor silently running on random weights is worse than an error
            from .._C import DroneYoloError

            fused = not any(".bn." in k for k in obj["model"])
            why = ("the checkpoint was saved after fuse() (no BatchNorm entries): load the unfused weights, fuse() runs here"
                   if fused else "the checkpoint belongs to another scale / YAML")
            raise DroneYoloError(f"{weights}: only {got} of {want} tensors match the model built from its YAML ({why}); first missing: {missing}")
        if "names" in obj:
            self.model.names = obj["names"]
        self.overrides = {"model": weights, "task": "detect"}

    @property
    def names(self):
        return self.model.names

    @property
    def device(self):
        return next(self.model.parameters()).device

    def fuse(self):
        self.model.fuse()
        return self

    def info(self, detailed=False, verbose=True):
        n_p = sum(p.numel() for p in self.model.parameters())
        n_l = len(list(self.model.modules()))
        if verbose:
            print(f"{self.model_name} summary: {n_l} layers, {n_p:,} parameters")
        return n_l, n_p

    def __call__(self, source=None, stream=False, **kwargs):
        return self.predict(source, stream, **kwargs)

    def predict(self, source=None, stream=False, predictor=None, **kwargs):
        """Run inference (engine/model.py:501-560): defaults conf=0.25, batch=1, mode=predict; the predictor is built
        on the first call and reused; a custom predictor class can be injected."""
        custom = {"conf": 0.25, "batch": 1, "mode": "predict"}
        args = {**self.overrides, **custom, **kwargs}
        args.pop("model", None)
        if args.get("augment"):
            from .._C import DroneYoloError

            raise DroneYoloError("augment=True (TTA) is outside the inference hot path")
        if self.predictor is None:
            self.predictor = (predictor or DetectionPredictor)(overrides=args)
            self.predictor.setup_model(model=self.model)
        else:  # only update args (engine/model.py:555-557); engines are keyed by the settings they depend on
            dev_changed = "device" in kwargs and str(kwargs["device"]) != str(getattr(self.predictor.args, "device", None))
            self.predictor.update_args(args)
            if dev_changed:
                self.predictor.setup_model(model=self.model)
        return self.predictor(source=source, stream=stream)


    def val(self, validator=None, batches=None, **kwargs):
        """Validate on an iterable of batch dicts (engine/model.py:596-626; datasets / dataloaders are outside the hot path, so the
        batches are passed in): returns the validator's metrics object (`box`-level numbers: mp, mr, map50, map, results_dict)."""
        from .validator import DetectionValidator

        args = {**{k: v for k, v in self.overrides.items() if k not in ("model", "task")}, **kwargs}
        v = (validator or DetectionValidator)(args=args)
        v(model=self.model, batches=batches)
        self.metrics = v.metrics
        return v.metrics


class YOLO(Model):
    """YOLO(model='yolov8s-p2-repvgg.yaml') — detection task only (models/yolo/model.py:35-40)."""

    @property
    def task_map(self):
        from .validator import DetectionValidator

        return {"detect": {"model": DetectionModel, "predictor": DetectionPredictor, "validator": DetectionValidator}}
