"""Writes tests/golden/ref_ckpt_tiny.pt: a checkpoint pickled by the REAL reference (torch.save of its DetectionModel, the
format `YOLO('x.pt')` / mix6.py:18 read), for a tiny width of the Drone-YOLO graph so that it stays small, plus
ref_ckpt_tiny.json with the digest of its state_dict.  Runs in the authoring container only (needs /root/reference)."""
import copy
import hashlib
import json
import sys
from pathlib import Path

import torch
import yaml

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import recipe, ref_shim  # noqa: E402

OUT = ROOT / "tests" / "golden"


def digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().float().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    tasks = ref_shim.load()
    cfg = yaml.safe_load((ref_shim.REF_ROOT / "ultralytics/cfg/models/v8/yolov8-p2-repvgg.yaml").read_text())
    cfg["scales"] = {"t": [0.33, 0.125, 1024]}             # half the n width: every channel count stays a multiple of 8
    cfg["scale"] = "t"
    cfg["nc"] = 10
    torch.manual_seed(0)
    model = tasks.DetectionModel(cfg, nc=10, verbose=False)
    recipe.apply_recipe(model)
    model.names = {i: f"cls{i}" for i in range(10)}
    ckpt = {"epoch": -1, "best_fitness": None, "model": copy.deepcopy(model).half(), "ema": None, "updates": None,
            "optimizer": None, "train_args": {"task": "detect", "imgsz": 640, "batch": 16}, "date": "fixture", "version": "8.3.82"}
    torch.save(ckpt, OUT / "ref_ckpt_tiny.pt")
    sd = {k: v.half().float() for k, v in model.state_dict().items()}
    (OUT / "ref_ckpt_tiny.json").write_text(json.dumps({"digest": digest(sd), "n_tensors": len(sd), "yaml_scale": "t",
                                                        "names": model.names}, indent=1))
    print((OUT / "ref_ckpt_tiny.pt").stat().st_size, "bytes,", len(sd), "tensors")


if __name__ == "__main__":
    main()
