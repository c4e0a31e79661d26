"""In-kernel timeline of one conv layer (debug build: DY_CONV_DEBUG_BUILD=1 python -m drone_yolo_b200.build -> lib/libdroneyolo_dbg.so, picked up automatically).

    python tools/trace_conv.py "P2 3x3 64->64" [more layer-name substrings]

Prints clock64() stamps of CTA 0 relative to its first event: A-producer, MMA and epilogue iterations.
"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
dev = torch.device("cuda:0")
trace = torch.zeros(5 * 96 * 8, dtype=torch.int64, device=dev)
os.environ["DY_CONV_TRACE"] = str(trace.data_ptr())
os.environ.setdefault("DY_LIB", str(ROOT / "drone_yolo_b200" / "lib" / "libdroneyolo_dbg.so"))
import tools.bench_conv as bc  # noqa: E402

ROLES = ["A-producer (0 pre-wait, 1 post-wait, 2 issued)", "MMA k-iter (0 pre-wait full, 1 post-wait, 2 issued, 3 committed)",
         "MMA tile (0 pre-wait tempty, 1 post-wait, 2 tile committed)",
         "epilogue group 0 (0 pre tfull, 1 post, 2 post tmem ld [+ residual wait], 3 post math+sts, 4 post barrier, 5 post store issue, 7 end)",
         "epilogue group 1"]
for pat in sys.argv[1:]:
    for layer in bc.LAYERS:
        if pat not in layer[0]:
            continue
        trace.zero_()
        bc.run(*layer, iters=1)
        t = trace.cpu().view(5, 96, 8)
        nz = t[t > 0]
        if nz.numel() == 0:
            print("no trace (release build?)")
            continue
        t0 = int(nz.min())
        print(f"=== {layer[0]}")
        for r in range(5):
            print(f"-- {ROLES[r]}")
            prev = None
            for i in range(96):
                row = t[r, i]
                if int(row.max()) == 0:
                    break
                if i < 40 or i % 8 == 0:
                    vals = [int(v) - t0 if int(v) > 0 else -1 for v in row]
                    d = "" if prev is None else f"  (+{vals[0] - prev})"
                    print(f"   {i:3d}: " + " ".join(f"{v:7d}" for v in vals if v >= 0) + d)
                prev = int(row[0]) - t0
