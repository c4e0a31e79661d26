"""Model construction for the Drone-YOLO graphs (reference ultralytics/nn/tasks.py): `yaml_model_load`,
`parse_model`, `BaseModel`, `DetectionModel`.

Same YAML grammar, same `model.<i>.*` state_dict keys, same seeded initialisation (module construction order is the
reference's, so `torch.manual_seed(s)` gives identical weights).  Differences, all on purpose:
  * `RepVGGBlock` is registered (the published reference tree lost that edit);
  * strides are derived from the graph instead of a 256x256 probe forward (tasks.py:323-337) — same values;
  * `fuse()` also re-parameterises RepVGGBlock (the reference skips it, SURVEY.md F5 — same math);
  * `forward` runs the layers through libdroneyolo's kernels; there is no PyTorch compute path.
"""
from __future__ import annotations

import ast
import contextlib
import copy
import math
import re
from pathlib import Path

import torch
import torch.nn as nn
import yaml

from .. import kernels as K
from .modules import C2f, Concat, Conv, Detect, DWConv, RepConv, RepVGGBlock, SPPF, Bottleneck, DFL

CFG_DIR = Path(__file__).resolve().parents[1] / "cfg"


def make_divisible(x, divisor):
    """Nearest multiple of divisor >= x (reference utils/ops.py:130-143)."""
    if isinstance(divisor, torch.Tensor):
        divisor = int(divisor.max())
    return math.ceil(x / divisor) * divisor


class Upsample(nn.Upsample):
    """nn.Upsample(None, 2, 'nearest') of the YAMLs, executed by dy_upsample2x."""

    def forward(self, x, out=None):
        if self.mode != "nearest" or float(self.scale_factor) != 2.0:
            raise K._C.DroneYoloError("Upsample: only scale_factor=2, mode='nearest' is implemented")
        from .modules.conv import as_input

        return K.upsample2x(as_input(x), out=out)


_MODULES = {
    "Conv": Conv, "DWConv": DWConv, "RepConv": RepConv, "RepVGGBlock": RepVGGBlock, "C2f": C2f, "SPPF": SPPF,
    "Concat": Concat, "Detect": Detect, "nn.Upsample": Upsample,
}
_BASE_MODULES = frozenset({Conv, DWConv, RepConv, RepVGGBlock, C2f, SPPF})   # take (c1, c2, *args)
_REPEAT_MODULES = frozenset({C2f})                                         # take n as an argument
_STRIDED = (Conv, DWConv, RepConv, RepVGGBlock)


def yaml_model_load(path):
    """Load a model YAML; 'yolov8s-p2-repvgg.yaml' resolves to yolov8-p2-repvgg.yaml with scale 's'
    (reference tasks.py:1093-1124)."""
    path = Path(path)
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", path.stem)
    candidates = [path, path.with_name(unified + path.suffix)]
    for name in (path.name, unified + path.suffix):
        candidates += list(CFG_DIR.rglob(name))
    for c in candidates:
        if c.is_file():
            d = yaml.safe_load(c.read_text())
            break
    else:
        raise FileNotFoundError(f"model YAML '{path}' not found (searched {CFG_DIR})")
    d["scale"] = guess_model_scale(path)
    d["yaml_file"] = str(path)
    return d


def guess_model_scale(model_path):
    """'yolov8s-p2.yaml' -> 's' (reference tasks.py:1109-1124)."""
    try:
        return re.search(r"yolo[v]?\d+([nslmx])", Path(model_path).stem).group(1)
    except AttributeError:
        return ""


def guess_model_task(model):
    """Only detection graphs exist here (reference tasks.py:1127-1190)."""
    return "detect"


def parse_model(d, ch, verbose=True):
    """YAML dict -> (nn.Sequential, save list), like reference tasks.py:929-1090 for the module set above.
    Also returns nothing else: per-layer reduction factors are stored on the modules as `.reduction`."""
    max_channels = float("inf")
    nc, scales = d.get("nc"), d.get("scales")
    depth, width = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0)
    if scales:
        scale = d.get("scale") or tuple(scales.keys())[0]
        depth, width, max_channels = scales[scale]
    Detect.legacy = True          # v8 graphs (reference tasks.py:933,1061-1062)
    ch = [ch]
    red = []                      # reduction factor (input px per output px) of every layer output
    layers, save, c2 = [], [], ch[-1]
    for i, (f, n, m, args) in enumerate(d["backbone"] + d["head"]):
        if m not in _MODULES:
            raise KeyError(f"module '{m}' is not part of the Drone-YOLO inference graphs ({sorted(_MODULES)})")
        m = _MODULES[m]
        args = list(args)
        for j, a in enumerate(args):
            if isinstance(a, str):
                with contextlib.suppress(ValueError, SyntaxError):
                    args[j] = nc if a == "nc" else ast.literal_eval(a)
        n = n_ = max(round(n * depth), 1) if n > 1 else n
        r_in = (red[f] if f != -1 else (red[-1] if red else 1)) if isinstance(f, int) else None
        if m in _BASE_MODULES:
            c1, c2 = ch[f], args[0]
            if c2 != nc:
                c2 = make_divisible(min(c2, max_channels) * width, 8)
            args = [c1, c2, *args[1:]]
            if m in _REPEAT_MODULES:
                args.insert(2, n)
                n = 1
            s = args[3] if m in _STRIDED and len(args) > 3 else 1
            r_out = r_in * s
        elif m is Concat:
            c2 = sum(ch[x] for x in f)
            rs = {red[x] for x in f}
            if len(rs) != 1:
                raise ValueError(f"layer {i}: Concat of maps with different resolutions {rs}")
            r_out = rs.pop()
        elif m is Detect:
            args.append([ch[x] for x in f])
            r_out = [red[x] for x in f]
        elif m is Upsample:
            c2 = ch[f]
            r_out = r_in / 2
        else:
            c2 = ch[f]
            r_out = r_in
        m_ = nn.Sequential(*(m(*args) for _ in range(n))) if n > 1 else m(*args)
        m_.np = sum(x.numel() for x in m_.parameters())
        m_.i, m_.f, m_.type = i, f, f"{m.__module__}.{m.__name__}".replace("drone_yolo_b200", "ultralytics")
        m_.reduction = r_out
        if verbose:
            print(f"{i:>3}{str(f):>20}{n_:>3}{m_.np:10.0f}  {m_.type:<45}{str(args):<30}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(m_)
        if i == 0:
            ch = []
        ch.append(c2)
        red.append(r_out)
    return nn.Sequential(*layers), sorted(save)


def initialize_weights(model):
    """BN eps / momentum and in-place activations (reference utils/torch_utils.py:423-433)."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.eps = 1e-3
            m.momentum = 0.03
        elif isinstance(m, (nn.Hardswish, nn.LeakyReLU, nn.ReLU, nn.ReLU6, nn.SiLU)):
            m.inplace = True


class BaseModel(nn.Module):
    """Sequential executor with a skip list (reference tasks.py:95-297)."""

    def forward(self, x, *args, **kwargs):
        return self.predict(x, *args, **kwargs)

    def predict(self, x, profile=False, visualize=False, augment=False, embed=None):
        if augment or visualize or embed or profile:
            raise K._C.DroneYoloError("augment / visualize / embed / profile are outside the inference hot path")
        return self._predict_once(x)

    def _predict_once(self, x, profile=False, visualize=False, embed=None):
        """Layer loop of reference tasks.py:134-161; every module launches libdroneyolo kernels."""
        y = []
        for m in self.model:
            if m.f != -1:
                x = y[m.f] if isinstance(m.f, int) else [x if j == -1 else y[j] for j in m.f]
            x = m(x)
            y.append(x if m.i in self.save else None)
        return x

    def fuse(self, verbose=True):
        """Fold every Conv+BN and re-parameterise RepConv / RepVGGBlock (reference tasks.py:193-221 plus the
        `switch_to_deploy` the reference forgets)."""
        if not self.is_fused():
            for m in self.model.modules():
                if isinstance(m, Conv) and hasattr(m, "bn"):
                    m.fuse()
                elif isinstance(m, RepConv):
                    m.fuse_convs()
                    m.forward = m.forward_fuse
                elif isinstance(m, RepVGGBlock):
                    m.switch_to_deploy()
            for m in self.model.modules():
                if hasattr(m, "invalidate"):
                    m.invalidate()
        return self

    def is_fused(self, thresh=10):
        """True when fewer than `thresh` BatchNorm layers remain (reference tasks.py:223-234)."""
        bn = tuple(v for k, v in nn.__dict__.items() if "Norm" in k)
        return sum(isinstance(v, bn) for v in self.modules()) < thresh

    def _apply(self, fn):
        """Move Detect's stride / anchor tensors with the model (reference tasks.py:247-263)."""
        self = super()._apply(fn)
        m = self.model[-1]
        if isinstance(m, Detect):
            m.stride = fn(m.stride)
            m.anchors = fn(m.anchors)
            m.strides = fn(m.strides)
        return self

    def load(self, weights, verbose=True):
        """Load a state_dict or a model carrying one (reference tasks.py:265-279): intersecting keys only."""
        sd = weights["model"] if isinstance(weights, dict) and "model" in weights else weights
        sd = sd.float().state_dict() if isinstance(sd, nn.Module) else sd
        own = self.state_dict()
        csd = {k: v for k, v in sd.items() if k in own and own[k].shape == v.shape}
        self.load_state_dict(csd, strict=False)
        for m in self.modules():
            if hasattr(m, "invalidate"):
                m.invalidate()
        self.transferred = (len(csd), len(own), [k for k in own if k not in csd][:8])   # callers decide how strict to be
        if verbose:
            print(f"Transferred {len(csd)}/{len(own)} items from pretrained weights")
        return self


class DetectionModel(BaseModel):
    """YOLO detection model built from a YAML (reference tasks.py:299-388)."""

    def __init__(self, cfg="yolov8n-p2-repvgg.yaml", ch=3, nc=None, verbose=True):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        self.yaml["ch"] = ch = self.yaml.get("ch", ch)
        if nc and nc != self.yaml["nc"]:
            if verbose:
                print(f"Overriding model.yaml nc={self.yaml['nc']} with nc={nc}")
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(copy.deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = self.yaml.get("inplace", True)
        self.end2end = False
        self.task = "detect"
        self.args = {}

        m = self.model[-1]
        if isinstance(m, Detect):
            m.inplace = self.inplace
            m.stride = torch.tensor([float(r) for r in m.reduction])   # == 256 / H_i of the reference's probe
            self.stride = m.stride
            m.bias_init()
        else:
            self.stride = torch.Tensor([32])
        # The reference's stride probe is a TRAINING-mode forward of zeros (tasks.py:335-337): every activation is
        # exactly 0, so each BatchNorm ends up with running_mean 0, running_var 0.9*1 + 0.1*0 and one tracked batch.
        # Reproduce that so that seeded construction gives the reference's state_dict bit for bit.
        for bn in self.modules():
            if isinstance(bn, nn.BatchNorm2d):
                bn.running_var.mul_(0.9)
                bn.num_batches_tracked.fill_(1)
        initialize_weights(self)

    def eval(self):
        return super().eval()


def attempt_load_state(model: BaseModel, path):
    """Load a `.pt` that holds a plain state_dict (or {'model': state_dict}).  Pickled reference checkpoints
    ({'model': <DetectionModel>, ...}) are read by nn/ckpt.py load_reference_checkpoint, which YOLO('<file>.pt') uses."""
    obj = torch.load(path, map_location="cpu", weights_only=True)
    return model.load(obj)


__all__ = ("yaml_model_load", "guess_model_scale", "guess_model_task", "parse_model", "BaseModel", "DetectionModel",
           "make_divisible", "initialize_weights", "Upsample", "attempt_load_state",
           "Conv", "DWConv", "RepConv", "RepVGGBlock", "C2f", "SPPF", "Concat", "Detect", "Bottleneck", "DFL")
