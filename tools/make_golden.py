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


def quantile_conf(y, rank=60):
    """A confidence threshold that lets about `rank` anchors of image 0 through (best-class score), as a short decimal."""
    best = np.sort(y[0, 4:].max(0))[::-1]
    return float(f"{float(best[min(rank, best.size - 1)]):.4g}")


def decode_nms_fixture(tasks, mu, tag, imgsz=128, B=2, nc=10, only=None):
    import torchvision
    from ultralytics.utils import ops

    det = tasks.Detect(nc=nc, ch=(16, 32, 64, 128))
    det.stride = torch.tensor([4.0, 8.0, 16.0, 32.0])
    det.eval()
    raw = recipe.synthetic_raw_maps(B, imgsz, nc, mu)
    with torch.no_grad():
        y = det._inference([r.clone() for r in raw])           # (B, 14, A)
    out = {"y": y.numpy().astype(np.float32), "mu": mu, "imgsz": imgsz, "B": B, "nc": nc, "raw_seed": 1234}
    # the predictor's form (iou 0.45) at a confidence that leaves a few dozen rows: with conf 0.25 these synthetic scores give none
    out["predict_conf"] = quantile_conf(out["y"])

    cases = {
        "default": dict(conf_thres=0.001, iou_thres=0.7, max_det=300),
        "multilabel": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),
        "agnostic": dict(conf_thres=0.001, iou_thres=0.5, max_det=100, agnostic=True),
        "classes": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, classes=[1, 3, 7]),
        "maxnms": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=200),
        "predict": dict(conf_thres=out["predict_conf"], iou_thres=0.45, max_det=300),
        # validator with a-priori labels (autolabelling, ops.py:272-277): rows (cls, cx, cy, w, h); image 1 has none
        "labels": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True, labels=LABELS),
    }
    real_nms = torchvision.ops.nms
    if only:
        cases = {k: v for k, v in cases.items() if k in only}
    for name, kw in cases.items():
        kept = []
        sizes = []

        def recording_nms(boxes, scores, thr, _kept=kept):
            i = real_nms(boxes, scores, thr)
            _kept.append(i.numpy().copy())
            sizes.append(int(boxes.shape[0]))
            return i

        torchvision.ops.nms = recording_nms
        try:
            kw = dict(kw)
            if "labels" in kw:
                kw["labels"] = [torch.tensor(lb, dtype=torch.float32).reshape(-1, 5) for lb in kw["labels"]]
            res = ops.non_max_suppression(y.clone(), max_time_img=1e9, **kw)
            # H1: assert the reference's UNSTABLE pre-sort did not matter for this vector
            if kw.get("max_nms") or max(sizes, default=0) >= 30000:      # the truncation (ops.py:301-302) was exercised
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
        print(tag, name, [int(r.shape[0]) for r in res], "nms inputs", sizes)
    np.savez_compressed(OUT / f"decode_nms_{tag}.npz", **out)


def predict_fixture(tasks, tag, shapes, imgsz=128, conf=0.001, iou=0.7, max_det=300):
    """The reference's own `YOLO.predict` (engine/model.py:501-560 -> engine/predictor.py:221-306 ->
    models/yolo/detect/predict.py:23-73) on raw uint8 BGR frames: stores what its preprocess handed to the model, what the
    model handed to NMS, and the final `boxes.data` per frame, so that letterbox, NMS and scale_boxes/clip_boxes can each be
    compared bit for bit without the conv stack in between."""
    import shutil
    import tempfile

    from ultralytics import YOLO
    from ultralytics.utils import ops

    tmp = Path(tempfile.mkdtemp(prefix="dygold"))
    src = Path(tasks.__file__).resolve().parents[1] / "cfg" / "models" / "v8" / "yolov8-p2-repvgg.yaml"
    txt = src.read_text().replace("nc: 80", "nc: 10")
    assert "nc: 10" in txt
    (tmp / "yolov8n-p2-repvgg.yaml").write_text(txt)
    torch.manual_seed(0)
    yolo = YOLO(str(tmp / "yolov8n-p2-repvgg.yaml"))
    recipe.apply_recipe(yolo.model)
    frames = recipe.predict_frames(shapes, 100)
    seen = {}
    real_nms = ops.non_max_suppression

    def rec_nms(prediction, *a, **k):
        seen["y"] = (prediction[0] if isinstance(prediction, (list, tuple)) else prediction).detach().clone()
        return real_nms(prediction, *a, **k)

    def hook(predictor):
        if getattr(predictor, "_dy_wrapped", False):
            return
        pre = predictor.preprocess

        def wrapped(im):
            out = pre(im)
            seen["im"] = out.detach().clone()
            return out

        predictor.preprocess = wrapped
        predictor._dy_wrapped = True

    yolo.add_callback("on_predict_start", hook)
    ops.non_max_suppression = rec_nms
    try:
        res = yolo.predict(source=[f.copy() for f in frames], imgsz=imgsz, conf=conf, iou=iou, max_det=max_det, device="cpu", verbose=False)
    finally:
        ops.non_max_suppression = real_nms
        shutil.rmtree(tmp, ignore_errors=True)
    im = seen["im"]
    im_u8 = (im * 255.0).round().to(torch.uint8)
    assert torch.equal(im_u8.float() / 255.0, im), "preprocess output is not uint8 / 255"
    out = {"shapes": np.asarray(shapes, np.int64), "frame_seed0": 100, "imgsz": imgsz, "conf": conf, "iou": iou, "max_det": max_det,
           "nc": 10, "im_u8": im_u8.numpy(), "y": seen["y"].numpy().astype(np.float32)}
    for b, r in enumerate(res):
        out[f"boxes{b}"] = r.boxes.data.numpy().astype(np.float32)
        out[f"orig_shape{b}"] = np.asarray(r.orig_shape, np.int64)
    np.savez_compressed(OUT / f"predict_{tag}.npz", **out)
    print("predict", tag, tuple(im.shape), [int(r.boxes.data.shape[0]) for r in res])


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
    # BASELINE config 4 at full size: 34 000 anchors, one image per regime (dense: 33 5xx candidates > max_nms, H1 asserted)
    for mu, tag in ((-11.0, "sparse"), (-10.0, "vallike"), (-7.5, "dense")):
        decode_nms_fixture(tasks, mu, tag + "_34k", imgsz=640, B=1, only=("default", "multilabel", "predict"))
    # BASELINE config 2's model at its real resolution (raw maps as fp16: tolerance is 2e-2)
    convstack_fixture(tasks, "yolov8s-p2-repvgg.yaml", 640, 1, "s_repvgg_640_big")
    # the predictor end to end: differently shaped frames (full canvas) and equally shaped ones (rect / auto letterbox)
    predict_fixture(tasks, "ragged", [(300, 420), (250, 333)])
    predict_fixture(tasks, "rect", [(360, 640), (360, 640)], conf=0.01)


if __name__ == "__main__":
    main()
