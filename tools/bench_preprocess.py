"""Predictor preprocess for a list of raw frames: host path (cv2 letterbox + stack + transpose, then H2D of the canvas) vs
GPU path (H2D of the raw frames, one dy_letterbox_u8 per frame straight into the engine's input batch).

    python tools/bench_preprocess.py [--frames 64] [--h 1080] [--w 1920] [--imgsz 640]
"""
import argparse
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from drone_yolo_b200 import YOLO  # noqa: E402
from oracle import recipe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--h", type=int, default=1080)
ap.add_argument("--w", type=int, default=1920)
ap.add_argument("--imgsz", type=int, default=640)
a = ap.parse_args()
torch.manual_seed(0)
model = YOLO("yolov8s-p2-repvgg.yaml", nc=10)
recipe.apply_recipe(model.model)
frames = [np.random.default_rng(i).integers(0, 256, (a.h, a.w, 3), dtype=np.uint8) for i in range(a.frames)]
for gpu in (False, True):
    model.predictor = None
    for it in range(3):
        t0 = time.perf_counter()
        res = model.predict(frames, imgsz=a.imgsz, conf=0.25, device="cuda:0", gpu_preprocess=gpu)
        dt = time.perf_counter() - t0
    sp = res[0].speed
    print(f"gpu_preprocess={gpu}: {a.frames} frames {a.h}x{a.w} -> {a.imgsz}: {a.frames / dt:8.1f} frames/s wall "
          f"(per frame: preprocess {sp['preprocess']:.3f} ms, inference {sp['inference']:.3f} ms, postprocess {sp['postprocess']:.3f} ms)")
