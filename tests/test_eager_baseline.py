"""The same-box eager bar (oracle/eager_gpu.py, bench.py's `gpu_eager_baseline`) must compute the reference path: its decode
and NMS restatements are checked here against the pinned numpy oracle (CPU, fp32)."""
import numpy as np
import torch

from drone_yolo_b200.nn.tasks import DetectionModel
from oracle import eager_gpu, nms_np, recipe, torch_ref


def _model():
    torch.manual_seed(0)
    m = DetectionModel("yolov8n-p2-repvgg.yaml", nc=10, verbose=False)
    recipe.apply_recipe(m, cls_delta=2.65)
    return m.eval()


def test_eager_forward_matches_oracle_decode():
    m = _model()
    x = recipe.images(2, 64, 64)
    y_ref, _ = torch_ref.forward(m, x)
    y = eager_gpu.forward(m, x).numpy()
    np.testing.assert_allclose(y, y_ref, rtol=1e-5, atol=1e-4)
    m2 = eager_gpu.prepare(m, "cpu", torch.float32, deploy=False)
    y2 = eager_gpu.forward(m2, x.contiguous(memory_format=torch.channels_last)).numpy()
    np.testing.assert_allclose(y2, y_ref, rtol=1e-3, atol=1e-3)
    m3 = eager_gpu.prepare(m, "cpu", torch.float32, deploy=True)
    y3 = eager_gpu.forward(m3, x).numpy()
    np.testing.assert_allclose(y3, y_ref, rtol=1e-3, atol=2e-3)


def test_eager_nms_matches_oracle():
    m = _model()
    x = recipe.images(2, 128, 128)
    y_ref, _ = torch_ref.forward(m, x)
    for kw in (dict(), dict(multi_label=True), dict(agnostic=True), dict(classes=[0, 3])):
        ref = nms_np.non_max_suppression(y_ref, conf_thres=0.001, iou_thres=0.7, max_det=300, **kw)
        out = eager_gpu.non_max_suppression(torch.from_numpy(y_ref), 0.001, 0.7, max_det=300, **kw)
        assert sum(r.shape[0] for r in ref) > 0
        for a, b in zip(out, ref):
            assert np.array_equal(a.numpy().view(np.uint32), b.view(np.uint32))
