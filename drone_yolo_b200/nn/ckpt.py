"""Ingestion of the reference's pickled `.pt` checkpoints without the reference package on the path.

Replaces `torch_safe_load` / `attempt_load_one_weight` (ultralytics/nn/tasks.py:786-926) for the Drone-YOLO detect models
(`mix6.py:18` loads `Drone-YOLO.pt` this way).  Such a file is `torch.save({'model': <DetectionModel>, 'ema': ..,
'train_args': .., ...})`: the pickle refers to `ultralytics.nn.tasks.DetectionModel`, `ultralytics.nn.modules.*`, ... by
name.  Here every `ultralytics.*` class resolves to a stand-in (an `nn.Module` subclass for modules, a plain object for
the rest), which is enough for the generic `nn.Module` machinery to rebuild the module tree and hand back its `yaml` dict,
`names` and `state_dict()`; the real model is then built from that YAML with this package's modules and loaded with the
weights.  Outside `ultralytics.*` only an EXACT set of (module, name) pairs resolves: torch's tensor / storage rebuild
helpers, `torch.nn.modules.*` classes that are `nn.Module` subclasses, a few plain containers (OrderedDict, numpy array
reconstruction, pathlib paths, Namespace).  Dotted names, functions and every other global are refused, so a crafted file
cannot reach `torch.hub.load`, `numpy.testing...runstring`, `types.FunctionType` and the like (the reference's plain
`torch.load` executes whatever the pickle names).
"""
from __future__ import annotations

import pickle
import types
from typing import Any

import torch
import torch.nn as nn

# Exact (module, name) pairs a reference checkpoint may name outside `ultralytics.*`.  A prefix rule ("anything under torch /
# numpy / types") is not a protection: `numpy.testing._private.utils.runstring`, `types.FunctionType`, `torch.hub.load`, ...
# all live under such prefixes and execute their arguments.  Everything here is a data constructor or a class whose
# construction runs no caller-supplied code.
_TORCH_STORAGES = ("FloatStorage", "HalfStorage", "BFloat16Storage", "DoubleStorage", "LongStorage", "IntStorage", "ShortStorage",
                   "CharStorage", "ByteStorage", "BoolStorage", "UntypedStorage")
_ALLOWED_GLOBALS = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"), ("torch._utils", "_rebuild_parameter_with_state"),
    ("torch", "Size"), ("torch", "device"), ("torch", "dtype"), ("torch", "Tensor"),
    ("torch.nn.parameter", "Parameter"), ("torch.serialization", "_get_layout"),
    ("numpy", "dtype"), ("numpy", "ndarray"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "scalar"),
    ("_codecs", "encode"),
    ("pathlib", "PosixPath"), ("pathlib", "PurePosixPath"), ("pathlib", "WindowsPath"), ("pathlib", "PureWindowsPath"), ("pathlib", "Path"),
    ("argparse", "Namespace"), ("types", "SimpleNamespace"),
} | {("torch", n) for n in _TORCH_STORAGES}
_TORCH_DTYPES = {"float32", "float16", "bfloat16", "float64", "int64", "int32", "int16", "int8", "uint8", "bool"}
_ALLOWED_BUILTINS = {"set", "frozenset", "dict", "list", "tuple", "int", "float", "bool", "str", "bytes", "bytearray", "complex",
                     "slice", "range", "object"}
_stubs: dict = {}


class _Plain:
    """Stand-in for non-module reference classes (IterableSimpleNamespace, losses, ...): keeps whatever state it is given."""

    def __new__(cls, *a, **k):
        return object.__new__(cls)

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)


def _stub(module: str, name: str):
    key = (module, name)
    if key not in _stubs:
        is_module = ".nn." in module + "." or module.endswith(".nn") or name.endswith(("Model", "Detect"))
        base = nn.Module if is_module else _Plain
        _stubs[key] = type(name, (base,), {"__module__": module, "_droneyolo_stub": True})
    return _stubs[key]


class _Unpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str) -> Any:
        if module == "ultralytics" or module.startswith("ultralytics."):
            return _stub(module, name)
        if module in ("builtins", "__builtin__") and name in _ALLOWED_BUILTINS:
            return super().find_class("builtins", name)
        if "." in name:                                       # dotted names walk attributes (pickle protocol 4): never needed here
            raise pickle.UnpicklingError(f"checkpoint refers to the dotted name {module}.{name}: refused")
        if (module, name) in _ALLOWED_GLOBALS or (module == "torch" and name in _TORCH_DTYPES):
            return super().find_class(module, name)
        if module.startswith("torch.nn.modules."):            # torch's own layer classes (Conv2d, BatchNorm2d, SiLU, ...): classes only
            obj = super().find_class(module, name)
            if isinstance(obj, type) and issubclass(obj, nn.Module):
                return obj
        raise pickle.UnpicklingError(f"checkpoint refers to {module}.{name}, which is outside ultralytics / torch: refused")


def _pickle_module():
    m = types.ModuleType("droneyolo_ckpt_pickle")
    m.Unpickler = _Unpickler
    m.load = lambda f, **kw: _Unpickler(f, **kw).load()
    m.__name__ = "pickle"
    return m


def load_reference_checkpoint(path: str):
    """Returns (yaml_dict, state_dict (fp32), names or None, train_args or None) of a pickled reference checkpoint."""
    with open(path, "rb") as f:
        ckpt = torch.load(f, map_location="cpu", pickle_module=_pickle_module(), weights_only=False)
    model = ckpt if isinstance(ckpt, nn.Module) else None
    train_args = None
    if isinstance(ckpt, dict):
        model = ckpt.get("ema") or ckpt.get("model")
        train_args = ckpt.get("train_args")
    if isinstance(model, nn.Module) and not hasattr(model, "yaml") and hasattr(model, "model") and isinstance(model.model, nn.Module):
        model = model.model                                  # a pickled YOLO wrapper (tasks.py:855-860)
    if not isinstance(model, nn.Module) or not isinstance(getattr(model, "yaml", None), dict):
        raise ValueError(f"{path}: no pickled detection model with a `yaml` dict found (keys: {list(ckpt) if isinstance(ckpt, dict) else type(ckpt)})")
    state = {k: (v.float() if v.is_floating_point() else v) for k, v in model.state_dict().items()}
    names = getattr(model, "names", None)
    return dict(model.yaml), state, names, (dict(train_args) if isinstance(train_args, dict) else train_args)
