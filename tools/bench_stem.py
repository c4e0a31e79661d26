"""Time the stem (model.0) alone: tensor-core kernel vs the CUDA-core fallback (DY_STEM_CUDA_CORES=1)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from drone_yolo_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, C = 64, 640, 640, 32
w = torch.randn(C, 27, device=dev) * 0.3
b = torch.randn(C, device=dev)
for dt in (torch.uint8, torch.float32):
    x = (torch.rand(B, 3, H, W, device=dev) * 255).to(dt) if dt == torch.uint8 else torch.rand(B, 3, H, W, device=dev)
    out = K.empty_nhwc(B, C, H // 2, W // 2, dev)
    for _ in range(3):
        K.stem_conv(x, w, b, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        K.stem_conv(x, w, b, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    byts = x.numel() * x.element_size() + out.numel() * 2
    print(f"stem {dt}: {us:.1f} us  {byts / us / 1e3:.0f} GB/s")
