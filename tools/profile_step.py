"""One step of the hot path between cudaProfilerStart/Stop, for `ncu --profile-from-start off` launch lists.

    python tools/profile_step.py [--scale s] [--imgsz 640] [--batch 64] [--micro-batch 8] [--steps 1]
"""
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
ap.add_argument("--micro-batch", type=int, default=8)
ap.add_argument("--steps", type=int, default=1)
a = ap.parse_args()
dev = torch.device("cuda:0")
model = build_model(a.scale).to(dev).fuse(verbose=False)
eng = Engine(model, a.batch, a.imgsz, dev, micro_batch=a.micro_batch, conf=0.001, iou=0.7, cuda_graph=False)
eng.images.copy_(recipe.images(a.batch, a.imgsz, a.imgsz).to(dev))
for _ in range(2):
    eng.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    eng.step()
e1.record()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print(f"step {e0.elapsed_time(e1) / a.steps:.3f} ms, {eng.launches_per_step} launches, micro-batch {eng.mb}")
