"""drone_yolo_b200 — B200-native (sm_100a) implementation of the Drone-YOLO inference hot path
(conv stack -> Detect decode -> NMS) behind the reference's ultralytics plugin surface.

    from drone_yolo_b200 import YOLO
    results = YOLO("yolov8s-p2-repvgg.yaml", nc=10).predict(images, conf=0.25, iou=0.7)

All compute runs in libdroneyolo.so (hand-written CUDA, C-ABI in include/droneyolo.h); there is no CPU or
PyTorch fallback.  (The directory is `drone_yolo_b200`, the importable spelling of "drone-yolo_b200".)
"""
__version__ = "0.1.0"

from .engine.model import YOLO  # noqa: E402,F401
from .nn.tasks import DetectionModel  # noqa: E402,F401
from .engine.slicer import InferenceSlicer  # noqa: E402,F401
