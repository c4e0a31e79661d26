"""Per-phase clocks of the NMS selection on BASELINE config 4 inputs (debug build: DY_CONV_DEBUG_BUILD=1 python -m
drone_yolo_b200.build; DY_LIB=drone_yolo_b200/lib/libdroneyolo_dbg.so python tools/trace_nms_cfg4.py [--batch 256])."""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from drone_yolo_b200 import _C  # noqa: E402
from drone_yolo_b200 import kernels as K  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--imgsz", type=int, default=640)
a = ap.parse_args()
dev = torch.device("cuda:0")
B, nc, no, ld = a.batch, 10, 74, 80
shapes = [(a.imgsz // s, a.imgsz // s) for s in (4, 8, 16, 32)]
A = sum(h * w for h, w in shapes)
lib = _C.lib()
names = ["select passes", "gather", "sort", "box load", "vs kept", "bitmask", "scan", "copy kept"]
for mu, regime in ((-11.0, "sparse"), (-10.0, "val-like"), (-7.5, "dense")):
    g = torch.Generator(device=dev).manual_seed(1234)
    lv = []
    for h, w in shapes:
        t = torch.zeros((B, h, w, ld), device=dev, dtype=torch.float32)
        t[..., :64] = 1.5 * torch.randn((B, h, w, 64), device=dev, generator=g)
        t[..., 64:no] = mu + 1.5 * torch.randn((B, h, w, nc), device=dev, generator=g)
        lv.append(t.permute(0, 3, 1, 2)[:, :no])
    y = K.detect_decode(lv, [4.0, 8.0, 16.0, 32.0], nc)
    nb = K.NmsBuffers(B, nc, A, 300, False, dev)
    n = min(B, 1024)
    buf = (C.c_ulonglong * (16 * n))()
    K.nms(y, 0.001, 0.7, max_det=300, bufs=nb)
    lib.dy_nms_trace_read(buf, n)
    K.nms(y, 0.001, 0.7, max_det=300, bufs=nb)
    lib.dy_nms_trace_read(buf, n)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(n, 16).astype(np.float64)
    tot = t[:, :8].sum(1)
    print(f"== {regime} B={B} A={A}: candidates/image {t[:, 11].mean():.0f}, super-rounds {t[:, 9].mean():.2f}, rounds {t[:, 10].mean():.2f}, clocks/image mean {tot.mean():.0f} max {tot.max():.0f}")
    for k, nm in enumerate(names):
        print(f"  {nm:14s} mean {t[:, k].mean():9.0f}  ({100 * t[:, k].mean() / tot.mean():5.1f} %)")
    print(f"  thread 0: pair-loop trips {t[:, 14].mean():.1f}, heavy trips {t[:, 12].mean():.2f}, x-overlapping columns per trip {t[:, 13].mean() / max(t[:, 14].mean(), 1):.2f}")
    del lv, y, nb
