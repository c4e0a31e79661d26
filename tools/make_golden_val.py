"""tests/golden/val_*.npz: the REAL reference's DetectionValidator (models/yolo/detect/val.py) run method by method on a
synthetic batch — preprocess -> model -> postprocess (NMS, multi_label, conf 0.001) -> update_metrics -> get_stats.

Run in the authoring container only (the reference tree does not travel):  python tools/make_golden_val.py
An untrained model has mAP ~ 0 against arbitrary labels, which would make an mAP check vacuous; so the labels are MADE from
the reference's own single-label detections of the same images (a seeded subset, boxes jittered, some classes flipped):
the correct matrix then has hits at every IoU threshold and misses, and mAP sits well inside (0, 1).
Stored: the images (uint8), labels, the reference's pre-NMS tensor y, its NMS rows, the stats (tp, conf, pred_cls,
target_cls), AP per class and the results dict; a second case with `save_hybrid=True` (labels appended by NMS).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import recipe, ref_shim  # noqa: E402

OUT = ROOT / "tests" / "golden"
YAML, IMGSZ, B, NC, CLS_DELTA, CLS_GAIN = "yolov8n-p2-repvgg.yaml", 128, 8, 10, 0.0, 40.0   # spread scores: AP sorts by score
ORI = (100, 128)                                 # original frames 100 x 128, letterboxed into 128 x 128: gain 1, pad (0, 14)
RATIO_PAD = ((1.0, 1.0), (0.0, 14.0))


def make_labels(tasks, model, img_u8):
    """Labels from the reference's own detections: per image the 10 best single-label rows at conf 0.25-quantile, jittered."""
    from ultralytics.utils import ops

    with torch.no_grad():
        y = model(img_u8.float() / 255)[0]
    det = ops.non_max_suppression(y.clone(), 0.001, 0.5, max_det=300)
    g = torch.Generator().manual_seed(7)
    cls, boxes, bidx = [], [], []
    for i, d in enumerate(det):
        take = d[torch.randperm(min(len(d), 40), generator=g)[:10]]
        for r in take:
            x1, y1, x2, y2 = r[:4].tolist()
            w, h = x2 - x1, y2 - y1
            jit = (torch.rand(4, generator=g) - 0.5) * torch.tensor([w, h, w, h]) * 0.25      # IoU with the row ~0.6 .. 1
            x1, y1, x2, y2 = (torch.tensor([x1, y1, x2, y2]) + jit).tolist()
            x1, x2 = sorted((min(max(x1, 0.0), IMGSZ), min(max(x2, 0.0), IMGSZ)))
            y1, y2 = sorted((min(max(y1, 14.0), IMGSZ - 14.0), min(max(y2, 14.0), IMGSZ - 14.0)))
            if x2 - x1 < 2 or y2 - y1 < 2:
                continue
            c = int(r[5]) if float(torch.rand(1, generator=g)) > 0.15 else int(torch.randint(0, NC, (1,), generator=g))
            cls.append([float(c)])
            boxes.append([(x1 + x2) / 2 / IMGSZ, (y1 + y2) / 2 / IMGSZ, (x2 - x1) / IMGSZ, (y2 - y1) / IMGSZ])
            bidx.append(float(i))
    return (torch.tensor(cls, dtype=torch.float32), torch.tensor(boxes, dtype=torch.float32), torch.tensor(bidx, dtype=torch.float32))


def run_case(tasks, model, img_u8, labels, save_hybrid, tag):
    from ultralytics.models.yolo.detect import DetectionValidator

    cls, boxes, bidx = labels
    tmp = Path(tempfile.mkdtemp(prefix="dyval"))
    v = DetectionValidator(save_dir=tmp, args=dict(task="detect", mode="val", conf=0.001, iou=0.7, max_det=300, plots=False,
                                                   save_hybrid=save_hybrid, imgsz=IMGSZ, batch=B, half=False, save_json=False))
    v.device = torch.device("cpu")
    v.data = {"val": "", "names": model.names, "nc": NC}
    v.training = False
    v.init_metrics(model)
    batch = {"img": img_u8.clone(), "cls": cls.clone(), "bboxes": boxes.clone(), "batch_idx": bidx.clone(),
             "ori_shape": [ORI] * B, "ratio_pad": [RATIO_PAD] * B, "im_file": [f"im{i}.jpg" for i in range(B)]}
    batch = v.preprocess(batch)
    with torch.no_grad():
        preds = model(batch["img"])
    y = preds[0].clone()
    dets = v.postprocess(preds)
    rows = [d.clone().numpy() for d in dets]
    v.update_metrics(dets, batch)
    res = v.get_stats()
    stats = {k: torch.cat(x, 0).cpu().numpy() for k, x in v.stats.items()}
    m = v.metrics.box
    np.savez_compressed(
        OUT / f"val_{tag}.npz", yaml=YAML, imgsz=IMGSZ, B=B, nc=NC, model_seed=0, bn_seed=1, cls_delta=CLS_DELTA, cls_gain=CLS_GAIN, save_hybrid=save_hybrid,
        img=img_u8.numpy(), cls=cls.numpy(), bboxes=boxes.numpy(), batch_idx=bidx.numpy(), ori_shape=np.array(ORI),
        ratio_pad=np.array([RATIO_PAD[0][0], RATIO_PAD[0][1], RATIO_PAD[1][0], RATIO_PAD[1][1]]),
        y=y.numpy().astype(np.float32), n_rows=np.array([len(r) for r in rows]), rows=np.concatenate(rows, 0),
        tp=stats["tp"], conf=stats["conf"], pred_cls=stats["pred_cls"], target_cls=stats["target_cls"],
        all_ap=np.asarray(m.all_ap), ap_class_index=np.asarray(m.ap_class_index), p=np.asarray(m.p), r=np.asarray(m.r),
        results=np.array([res[k] for k in ("metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness")]))
    print(tag, "rows", [len(r) for r in rows], "labels", len(cls), "tp@.5", int(stats["tp"][:, 0].sum()), "tp@.95", int(stats["tp"][:, 9].sum()),
          {k: round(float(x), 4) for k, x in res.items()}, "distinct conf", len(np.unique(stats["conf"])), "of", len(stats["conf"]))


def main():
    tasks = ref_shim.load()
    torch.manual_seed(0)
    model = tasks.DetectionModel(YAML, nc=NC, verbose=False)
    recipe.apply_recipe(model, cls_delta=CLS_DELTA, cls_gain=CLS_GAIN)
    model.eval()
    model.names = {i: f"cls{i}" for i in range(NC)}
    model.fuse(verbose=False)
    img_u8 = (recipe.images(B, IMGSZ, IMGSZ, seed=5) * 255).round().to(torch.uint8)
    labels = make_labels(tasks, model, img_u8)
    run_case(tasks, model, img_u8, labels, False, "n128")
    run_case(tasks, model, img_u8, labels, True, "n128_hybrid")


if __name__ == "__main__":
    main()
