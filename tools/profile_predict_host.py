"""Where the HOST time of YOLO.predict(stream=True) goes when one process turns many images per engine step into Results
(rank 0 of an 8-GPU run builds 512 per step): cProfile over a few batches of `--batch` frames on one GPU.

    python tools/profile_predict_host.py --batch 512 --batches 6"""
import argparse
import cProfile
import pstats
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from bench import build_model  # noqa: E402
from drone_yolo_b200 import YOLO  # noqa: E402
from oracle import recipe  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--batches", type=int, default=6)
a = ap.parse_args()
dev = torch.device("cuda:0")
model = build_model("s").to(dev).fuse(verbose=False)
host8 = (recipe.images(64, 640, 640) * 255).round().to(torch.uint8)
frames_t = host8.permute(0, 2, 3, 1).flip(-1).contiguous().pin_memory()
frames = [f.numpy() for f in frames_t] * (a.batch // 64)
yolo = YOLO(model)
kw = dict(imgsz=640, conf=0.001, iou=0.7, max_det=300, device=dev, batch=a.batch, stream=True)


def stream(nb):
    for _ in range(nb):
        yield from frames


def run(nb):
    n = 0
    for r in yolo.predict(stream(nb), **kw):
        n += len(r)
    return n


run(3)
torch.cuda.synchronize()
t0 = time.perf_counter()
pr = cProfile.Profile()
pr.enable()
run(a.batches)
pr.disable()
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f"{a.batches} batches of {a.batch}: {dt / a.batches * 1e3:.2f} ms per batch wall (under cProfile), {a.batch * a.batches / dt:.0f} images/s")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(22)
