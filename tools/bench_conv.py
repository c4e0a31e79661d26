"""Time single conv layers (CUDA events) — used with the DY_CONV_DBG knock-outs to find what bounds the kernel.

    python tools/bench_conv.py            # a fixed list of representative Drone-YOLO-s layers at batch 64
"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from drone_yolo_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda:0")
LAYERS = [  # name, B, cin, cout, H, W, k, s, res, in_ld, out_ld, f32
    ("P2 1x1 64->64", 64, 64, 64, 160, 160, 1, 1, False, 64, 96, False),
    ("P2 1x1 192->64", 64, 192, 64, 160, 160, 1, 1, False, 192, 96, False),
    ("P2 3x3 32->32 +res", 64, 32, 32, 160, 160, 3, 1, True, 96, 96, False),
    ("P2 3x3 32->32 nores", 64, 32, 32, 160, 160, 3, 1, False, 96, 96, False),
    ("P2 3x3 32->32 dense", 64, 32, 32, 160, 160, 3, 1, False, 32, 32, False),
    ("P2 3x3 32->32 out_ld64", 64, 32, 32, 160, 160, 3, 1, False, 96, 64, False),
    ("P2 3x3 32->32 out_ld128", 64, 32, 32, 160, 160, 3, 1, False, 96, 128, False),
    ("P2 3x3 32->32 in32 out96", 64, 32, 32, 160, 160, 3, 1, False, 32, 96, False),
    ("P2 1x1 64->64 out_ld64", 64, 64, 64, 160, 160, 1, 1, False, 64, 64, False),
    ("P2 1x1 64->64 out_ld128", 64, 64, 64, 160, 160, 1, 1, False, 64, 128, False),
    ("P2 3x3 64->64", 64, 64, 64, 160, 160, 3, 1, False, 128, 128, False),
    ("P3 3x3 64->64", 64, 64, 64, 80, 80, 3, 1, False, 128, 128, False),
    ("P2 3x3 64->128", 64, 64, 128, 160, 160, 3, 1, False, 64, 128, False),
    ("P2 1x1 64->64 f32", 64, 64, 64, 160, 160, 1, 1, False, 128, 80, True),
    ("P1->P2 3x3 s2 32->64", 64, 32, 64, 320, 320, 3, 2, False, 32, 64, False),
    ("P3 3x3 128->128", 64, 128, 128, 80, 80, 3, 1, False, 128, 128, False),
    ("P4 3x3 128->128", 64, 128, 128, 40, 40, 3, 1, False, 128, 128, False),
    ("P5 3x3 256->256", 64, 256, 256, 20, 20, 3, 1, False, 256, 256, False),
    ("P5 1x1 1024->512", 64, 1024, 512, 20, 20, 1, 1, False, 1024, 512, False),
    ("P3 3x3 64->64 +res", 64, 64, 64, 80, 80, 3, 1, True, 192, 192, False),
    ("P4 3x3 128->128 +res", 64, 128, 128, 40, 40, 3, 1, True, 384, 384, False),
    ("P3 1x1 256->128", 64, 256, 128, 80, 80, 1, 1, False, 256, 128, False),
    ("P4 1x1 512->256", 64, 512, 256, 40, 40, 1, 1, False, 512, 256, False),
    ("P3 3x3 s2 128->256", 64, 128, 256, 80, 80, 3, 2, False, 128, 256, False),
    ("P4 3x3 256->128", 64, 256, 128, 40, 40, 3, 1, False, 256, 128, False),
    ("P2->P3 3x3 s2 64->128", 64, 64, 128, 160, 160, 3, 2, False, 64, 128, False),
    ("P5 1x1 512->512", 64, 512, 512, 20, 20, 1, 1, False, 512, 768, False),
    ("P5 1x1 768->512", 64, 768, 512, 20, 20, 1, 1, False, 768, 512, False),
    ("P4 3x3 s2 256->512", 64, 256, 512, 40, 40, 3, 2, False, 256, 512, False),
    ("P4 1x1 384->256", 64, 384, 256, 40, 40, 1, 1, False, 384, 256, False),
]


def run(name, B, cin, cout, H, W, k, s, res, in_ld, out_ld, f32, iters=10):
    xb = torch.randn(B, H, W, in_ld, device=dev).to(torch.bfloat16)
    x = xb.permute(0, 3, 1, 2)[:, :cin]
    w = torch.randn(cout, cin, k, k, device=dev) / (cin * k * k) ** 0.5
    wp, bp = K.pack_conv_weight(w, torch.randn(cout, device=dev))
    Ho, Wo = H // s, W // s
    ob = torch.empty(B, Ho, Wo, out_ld, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    out = ob.permute(0, 3, 1, 2)[:, :cout]
    r = ob.permute(0, 3, 1, 2)[:, out_ld - cout:] if res else None      # another slice of the same buffer, like C2f
    if res and out_ld - cout < cout:
        r = torch.randn(B, Ho, Wo, cout, device=dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    for _ in range(2):
        K.conv2d(x, wp, bp, cout, k, s, True, residual=r, out=out)
    torch.cuda.synchronize()
    # the launches are replayed from a CUDA graph (as the engine does): issued from Python, every call re-encodes its tensor
    # maps on the host (~20 us) and the GPU idles between launches, which hides everything below that
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters):
                K.conv2d(x, wp, bp, cout, k, s, True, residual=r, out=out)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    reps = 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (iters * reps)
    M = B * Ho * Wo
    flop = 2.0 * M * cout * cin * k * k
    byts = B * H * W * cin * 2 + M * cout * (4 if f32 else 2) + (M * cout * 2 if res else 0)
    print(f"{name:24s} {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s  {byts / us / 1e3:7.1f} GB/s (algorithmic)", flush=True)


if __name__ == "__main__":
    import os
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    LAYERS = [l for l in LAYERS if only in l[0]]
    print("DY_CONV_DBG =", os.environ.get("DY_CONV_DBG", "0"))
    for layer in LAYERS:
        run(*layer)
