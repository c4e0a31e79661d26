"""Warm per-layer device times of one engine configuration (Engine.profile -> dy_program_profile), one line per op.
    python tools/profile_layers.py [--scale s] [--imgsz 640] [--batch 64] [--reps 5]"""
import argparse
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import build_model  # noqa: E402
from drone_yolo_b200.engine.engine import Engine  # noqa: E402
from oracle import recipe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scale", default="s")
ap.add_argument("--imgsz", type=int, default=640)
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda:0")
eng = Engine(build_model(a.scale).to(dev).fuse(verbose=False), a.batch, a.imgsz, dev, micro_batch=a.batch, conf=0.001, iou=0.7, cuda_graph=False)
eng.images.copy_(recipe.images(a.batch, a.imgsz, a.imgsz).to(dev))
for _ in range(3):
    eng.step()
torch.cuda.synchronize()
eng.profile(reps=a.reps, verbose=True)
