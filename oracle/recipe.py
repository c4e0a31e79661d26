"""Seeded test-weight / test-input recipes shared by tools/make_golden.py (run on the REAL reference) and by the
tests and bench (run on drone_yolo_b200 models).  TEST INFRASTRUCTURE.

Default-initialised weights are degenerate for parity work: BN statistics are (0, 0.9) so folding is a
near-identity, and the class bias keeps every score below conf=0.001 except on the stride-32 level.  The recipe
randomises every BatchNorm and shifts the class bias so that a realistic fraction of anchors survives."""
import hashlib

import numpy as np
import torch
import torch.nn as nn


def randomize_bn(model: nn.Module, seed: int = 1) -> None:
    """running_mean ~ 0.1 N(0,1), running_var ~ U(0.5,1.5), weight ~ U(0.5,1.5), bias ~ 0.1 N(0,1), in module order."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, nn.BatchNorm2d):
                n = m.num_features
                m.running_mean.copy_(0.1 * torch.randn(n, generator=g))
                m.running_var.copy_(0.5 + torch.rand(n, generator=g))
                m.weight.copy_(0.5 + torch.rand(n, generator=g))
                m.bias.copy_(0.1 * torch.randn(n, generator=g))


def shift_cls_bias(model: nn.Module, delta: float = 4.0) -> None:
    """Add `delta` to the class-logit bias of every Detect level (model.model[-1].cv3[l][-1].bias)."""
    det = model.model[-1]
    with torch.no_grad():
        for seq in det.cv3:
            seq[-1].bias.add_(delta)


def scale_cls_weight(model: nn.Module, gain: float) -> None:
    """Multiply the class-logit weights of every Detect level (model.model[-1].cv3[l][-1].weight) by `gain`: at the default
    initialisation the class logits hardly vary over the anchors, so the scores tie by the hundred; metrics that sort by score
    (AP) then depend on the order inside ties.  A gain of a few spreads the scores."""
    det = model.model[-1]
    with torch.no_grad():
        for seq in det.cv3:
            seq[-1].weight.mul_(gain)


def apply_recipe(model: nn.Module, bn_seed: int = 1, cls_delta: float = 4.0, cls_gain: float = 1.0) -> nn.Module:
    randomize_bn(model, bn_seed)
    shift_cls_bias(model, cls_delta)
    if cls_gain != 1.0:
        scale_cls_weight(model, cls_gain)
    for m in model.modules():
        if hasattr(m, "invalidate"):
            m.invalidate()
    return model


def images(B: int, H: int, W: int, seed: int = 2) -> torch.Tensor:
    return torch.rand(B, 3, H, W, generator=torch.Generator().manual_seed(seed))


def state_digest(model: nn.Module) -> str:
    """sha256 over every state_dict tensor (key order, raw bytes)."""
    h = hashlib.sha256()
    for k, v in model.state_dict().items():
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def level_shapes(imgsz: int):
    return [(imgsz // s, imgsz // s) for s in (4, 8, 16, 32)]


def synthetic_raw_maps(B: int, imgsz: int, nc: int, mu: float, seed: int = 1234):
    """BASELINE config 4 inputs: box logits 1.5 N(0,1), class logits N(mu, 1.5^2); per level (B, 64+nc, H, W) fp32.
    mu = -11 sparse (~3 % of anchors pass conf .001), -10 val-like (~18 %), -7.5 dense (~99 %)."""
    g = torch.Generator().manual_seed(seed)
    maps = []
    for h, w in level_shapes(imgsz):
        box = 1.5 * torch.randn(B, 64, h, w, generator=g)
        cls = mu + 1.5 * torch.randn(B, nc, h, w, generator=g)
        maps.append(torch.cat((box, cls), 1).contiguous())
    return maps


def ties_free(scores: np.ndarray) -> bool:
    return np.unique(scores).size == scores.size


def predict_frames(shapes, seed0: int = 100):
    """Raw uint8 BGR frames of the predict-level fixtures (tools/make_golden.py predict_fixture): seeded noise over a
    low-frequency pattern, so that the letterbox resize has something to interpolate."""
    frames = []
    for i, (h, w) in enumerate(shapes):
        f = np.random.RandomState(seed0 + i).randint(0, 256, (h, w, 3), dtype=np.uint8)
        yy, xx = np.mgrid[0:h, 0:w]
        ramp = ((np.sin(xx / 17.0 + i) + np.cos(yy / 23.0)) * 50 + 128).clip(0, 255).astype(np.uint8)
        f[...] = (f.astype(np.uint16) // 4 + ramp[..., None].astype(np.uint16) * 3 // 4).astype(np.uint8)
        frames.append(f)
    return frames
