"""Where the NMS select kernel spends its clocks (debug build of the library: DY_CONV_DEBUG_BUILD=1 python -m
drone_yolo_b200.build, then DY_LIB=drone_yolo_b200/lib/libdroneyolo_dbg.so python tools/trace_nms.py).

Runs the bench workload's conv stack once, then NMS on its prediction tensor, and prints per-phase clock totals of
nms_select_kernel (thread 0 of every image's CTA, barrier to barrier)."""
import argparse
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import build_model  # noqa: E402
from drone_yolo_b200 import _C  # noqa: E402
from drone_yolo_b200.engine.engine import Engine  # noqa: E402
from oracle import recipe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--cls-delta", type=float, default=2.5)
a = ap.parse_args()
dev = torch.device("cuda:0")
model = build_model("s", a.cls_delta).to(dev).fuse(verbose=False)
eng = Engine(model, a.batch, 640, dev, micro_batch=a.batch, conf=0.001, iou=0.7, cuda_graph=False)
eng.images.copy_(recipe.images(a.batch, 640, 640).to(dev))
lib = _C.lib()
buf = (C.c_ulonglong * (16 * a.batch))()
eng.step()
lib.dy_nms_trace_read(buf, a.batch)          # clear
eng.step()
lib.dy_nms_trace_read(buf, a.batch)
t = np.frombuffer(buf, dtype=np.uint64).reshape(a.batch, 16).astype(np.float64)
names = ["select passes", "gather", "sort", "box load", "vs kept", "bitmask", "scan", "copy kept"]
print(f"candidates/image: mean {t[:, 11].mean():.0f}  min {t[:, 11].min():.0f}  max {t[:, 11].max():.0f}")
print(f"super-rounds/image: mean {t[:, 9].mean():.2f} max {t[:, 9].max():.0f}; rounds/image: mean {t[:, 10].mean():.2f} max {t[:, 10].max():.0f}")
tot = t[:, :8].sum(1)
print(f"clocks/image: mean {tot.mean():.0f} max {tot.max():.0f}")
for k, nm in enumerate(names):
    print(f"  {nm:14s} mean {t[:, k].mean():9.0f}  ({100 * t[:, k].mean() / tot.mean():5.1f} %)   slowest image {t[int(tot.argmax()), k]:9.0f}")
