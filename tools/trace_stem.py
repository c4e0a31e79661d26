"""In-kernel timeline of the tensor-core stem (debug build: DY_CONV_DEBUG_BUILD=1 python -m drone_yolo_b200.build).

    python tools/trace_stem.py

Prints clock64() stamps of CTA 0 relative to its first event, per role and tile (each role sees every second tile)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
dev = torch.device("cuda:0")
trace = torch.zeros(9 * 96 * 8, dtype=torch.int64, device=dev)
os.environ["DY_CONV_TRACE"] = str(trace.data_ptr())
os.environ.setdefault("DY_LIB", str(ROOT / "drone_yolo_b200" / "lib" / "libdroneyolo_dbg.so"))
from drone_yolo_b200 import kernels as K  # noqa: E402

ROLES = ["producer 0 (0 pre-wait pempty, 1 post, 2 issued)", "producer 1",
         "MMA 0 (0 pre tempty, 1 post, 2 post afull, 3 committed)", "MMA 1",
         "builder 0 (0 pre pfull, 1 post, 2 values ready, 3 post aempty, 4 stored+arrived)", "builder 1",
         "epilogue 0 (1 bias ready, 2 post tmem ld wait, 3 post math+sts, 4 next tmem ld issued, 5 post barrier, 6 store issued)", "epilogue 1", "epilogue 2"]
B, H, W, C = 64, 640, 640, 32
w = torch.randn(C, 27, device=dev) * 0.3
b = torch.randn(C, device=dev)
x = (torch.rand(B, 3, H, W, device=dev) * 255).to(torch.uint8)
out = K.empty_nhwc(B, C, H // 2, W // 2, dev)
for _ in range(2):
    K.stem_conv(x, w, b, out=out)
torch.cuda.synchronize()
trace.zero_()
K.stem_conv(x, w, b, out=out)
torch.cuda.synchronize()
t = trace.cpu().view(9, 96, 8)
nz = t[t > 0]
if nz.numel() == 0:
    print("no trace (release build?)")
    sys.exit(0)
t0 = int(nz.min())
for r in range(9):
    print(f"-- {ROLES[r]}")
    prev = None
    for i in range(96):
        row = t[r, i]
        if int(row.max()) == 0:
            break
        if i < 24 or i % 8 == 0:
            vals = [int(v) - t0 if int(v) > 0 else -1 for v in row]
            d = "" if prev is None else f"  (+{vals[0] - prev})"
            print(f"   {i:3d}: " + " ".join(f"{v:7d}" for v in vals if v >= 0) + d)
        prev = int(row[0]) - t0
