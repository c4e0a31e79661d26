"""Generate tests/golden/*.npz by running the REAL reference (/root/reference, ultralytics fork) on CPU.

Run in the authoring container only (the reference tree does not travel to the GPU box):
    python tools/make_golden.py
Every fixture stores the reference's OUTPUTS plus the seeds that regenerate the inputs; inputs that cannot be
regenerated bit-for-bit without the reference (the pre-NMS tensors) are stored too.
"""
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import recipe, ref_shim  # noqa: E402

OUT = ROOT / "tests" / "golden"


def ref_model(tasks, yaml_name, nc=10, seed=0):
    torch.manual_seed(seed)
    m = tasks.DetectionModel(yaml_name, nc=nc, verbose=False)
    recipe.apply_recipe(m)
    return m.eval()


def convstack_fixture(tasks, yaml_name, imgsz, B, tag):
    m = ref_model(tasks, yaml_name)
    digest_unfused = recipe.state_digest(m)
    keys = list(m.state_dict().keys())
    n_params = sum(p.numel() for p in m.parameters())
    x = recipe.images(B, imgsz, imgsz)
    m.fuse(verbose=False)                       # what AutoBackend does before predict (autobackend.py:151-152)
    with torch.no_grad():
        y, raw = m(x)
    np.savez_compressed(
        OUT / f"convstack_{tag}.npz", yaml=yaml_name, imgsz=imgsz, B=B, nc=10, model_seed=0, bn_seed=1, cls_delta=4.0,
        image_seed=2, state_digest=digest_unfused, n_keys=len(keys), stride=m.stride.numpy(),
        n_params=n_params,
        y=y.numpy().astype(np.float32), **{f"raw{i}": r.numpy().astype(np.float16 if tag.endswith("big") else np.float32)
                                           for i, r in enumerate(raw)})
    print(tag, "y", tuple(y.shape), "digest", digest_unfused[:12])


LABELS = [[[3.0, 40.0, 52.0, 30.0, 22.0], [7.0, 90.5, 30.25, 12.0, 44.0], [3.0, 41.0, 51.0, 28.0, 24.0]], []]


def decode_nms_fixture(tasks, mu, tag, imgsz=128, B=2, nc=10):
    import torchvision
    from ultralytics.utils import ops

    det = tasks.Detect(nc=nc, ch=(16, 32, 64, 128))
    det.stride = torch.tensor([4.0, 8.0, 16.0, 32.0])
    det.eval()
    raw = recipe.synthetic_raw_maps(B, imgsz, nc, mu)
    with torch.no_grad():
        y = det._inference([r.clone() for r in raw])           # (B, 14, A)
    out = {"y": y.numpy().astype(np.float32), "mu": mu, "imgsz": imgsz, "B": B, "nc": nc, "raw_seed": 1234}

    cases = {
        "default": dict(conf_thres=0.001, iou_thres=0.7, max_det=300),
        "multilabel": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),
        "agnostic": dict(conf_thres=0.001, iou_thres=0.5, max_det=100, agnostic=True),
        "classes": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, classes=[1, 3, 7]),
        "maxnms": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=200),
        "predict": dict(conf_thres=0.25, iou_thres=0.45, max_det=300),
        # validator with a-priori labels (autolabelling, ops.py:272-277): rows (cls, cx, cy, w, h); image 1 has none
        "labels": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True, labels=LABELS),
    }
    real_nms = torchvision.ops.nms
    for name, kw in cases.items():
        kept = []

        def recording_nms(boxes, scores, thr, _kept=kept):
            i = real_nms(boxes, scores, thr)
            _kept.append(i.numpy().copy())
            return i

        torchvision.ops.nms = recording_nms
        try:
            kw = dict(kw)
            if "labels" in kw:
                kw["labels"] = [torch.tensor(lb, dtype=torch.float32).reshape(-1, 5) for lb in kw["labels"]]
            res = ops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
            # H1: assert the reference's UNSTABLE pre-sort did not matter for this vector
            if kw.get("max_nms"):
                orig_argsort = torch.Tensor.argsort
                torch.Tensor.argsort = lambda self, *a, **k: orig_argsort(self, *a, stable=True, **k)
                try:
                    res2 = ops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
                finally:
                    torch.Tensor.argsort = orig_argsort
                assert all(torch.equal(a, b) for a, b in zip(res, res2)), "unstable != stable pre-sort; change the seed"
        finally:
            torchvision.ops.nms = real_nms
        # images with no candidates never reach torchvision: align kept lists with images via the row counts
        it = iter(kept)
        for b, r in enumerate(res):
            out[f"{name}_out{b}"] = r.numpy().astype(np.float32)
            k = next(it)[: kw["max_det"]] if r.shape[0] or False else None
            if k is None:
                # zero rows: either no candidates (nms not called) or nms returned nothing (impossible: n>0 keeps >=1)
                k = np.zeros((0,), np.int64)
            out[f"{name}_kept{b}"] = k.astype(np.int64)
        print(tag, name, [int(r.shape[0]) for r in res])
    np.savez_compressed(OUT / f"decode_nms_{tag}.npz", **out)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    tasks = ref_shim.load()
    torch.set_num_threads(max(1, (os.cpu_count() or 2) - 1))
    convstack_fixture(tasks, "yolov8n-p2-repvgg.yaml", 128, 2, "n_repvgg_128")
    convstack_fixture(tasks, "yolov8n-p2-repvgg-sf.yaml", 64, 1, "n_repvgg_sf_64")
    convstack_fixture(tasks, "yolov8n-p2.yaml", 64, 1, "n_p2_64")
    convstack_fixture(tasks, "yolov8s-p2-repvgg.yaml", 64, 1, "s_repvgg_64")
    for mu, tag in ((-11.0, "sparse"), (-10.0, "vallike"), (-7.5, "dense")):
        decode_nms_fixture(tasks, mu, tag)


if __name__ == "__main__":
    main()
