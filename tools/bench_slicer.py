"""Tiled 4K frame (mix6.py:84-89: slice 2160x2160, overlap 0.2 -> 6 ragged tiles): one predict call per tile (the reference's
InferenceSlicer callback flow, here with this repo's predictor) against ONE batch through engine.slicer.InferenceSlicer.
Wall clock per frame through the public API, host buffers in, merged detections out.  Prints one JSON line."""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from drone_yolo_b200 import YOLO, InferenceSlicer  # noqa: E402
from drone_yolo_b200 import kernels as K  # noqa: E402
from drone_yolo_b200.engine.slicer import generate_offsets  # noqa: E402
from oracle import recipe  # noqa: E402  (seeded weights only)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", default="s")
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--hw", type=int, nargs=2, default=(2160, 3840))
    ap.add_argument("--conf", type=float, default=0.25, help="0.001: every tile returns max_det rows, the merge sees 1800 rows")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = YOLO(f"yolov8{a.scale}-p2-repvgg.yaml", nc=10)
    recipe.apply_recipe(model.model)
    frames = [np.random.default_rng(i).integers(0, 256, (*a.hw, 3), dtype=np.uint8) for i in range(3)]
    kw = dict(imgsz=640, conf=a.conf, iou=0.7, device=dev)
    offsets = generate_offsets((a.hw[1], a.hw[0]), (2160, 2160), (0.2, 0.2))

    def per_tile(frame):
        rows = []
        for x0, y0, x1, y1 in offsets.tolist():
            r = model.predict(frame[y0:y1, x0:x1], **kw)[0]
            d = np.asarray(r.boxes.data, dtype=np.float64)
            d[:, [0, 2]] += x0
            d[:, [1, 3]] += y0
            rows.append(d)
        rows = np.concatenate(rows, 0)
        if len(rows):
            rows = rows[K.box_nms_f64(torch.from_numpy(rows).to(dev), 0.7).cpu().numpy()]
        return rows

    slicer = InferenceSlicer(model, slice_wh=(2160, 2160), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.7, **kw)
    out = {"frame_hw": list(a.hw), "tiles": len(offsets), "scale": a.scale, "conf": a.conf}
    for name, fn in (("per_tile_calls", per_tile), ("one_batch", slicer)):
        model.predictor = None
        for i in range(3):
            fn(frames[i % 3])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(a.frames):
            fn(frames[i % 3])
        torch.cuda.synchronize()
        out[name + "_ms_per_frame"] = (time.perf_counter() - t0) * 1e3 / a.frames
    last = slicer(frames[0])
    out["merged_rows"] = len(last)
    out["speedup"] = out["per_tile_calls_ms_per_frame"] / out["one_batch_ms_per_frame"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
