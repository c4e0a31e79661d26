"""CPU tests of the host side: C-ABI symbol table, model construction for every scale / graph, re-parameterisation
algebra, box helpers, the loud failure without CUDA, and the 2-rank (gloo) shard + gather logic."""
import ctypes
import os
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]

PARAMS = {  # BASELINE.md §2 (nc = 10)
    "yolov8n-p2-repvgg.yaml": 2972360, "yolov8s-p2-repvgg.yaml": 10815576, "yolov8m-p2-repvgg.yaml": 25375784,
    "yolov8l-p2-repvgg.yaml": 43288952, "yolov8x-p2-repvgg.yaml": 67274376,
    "yolov8n-p2-repvgg-sf.yaml": 2978856, "yolov8s-p2-repvgg-sf.yaml": 10839320, "yolov8x-p2-repvgg-sf.yaml": 67414376,
}


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    from drone_yolo_b200 import _C

    g.build()
    header = (ROOT / "include" / "droneyolo.h").read_text()
    declared = set(re.findall(r"\b(dy_[a-z0-9_]+)\s*\(", header))
    declared -= {"dy_status"}
    lib = ctypes.CDLL(str(_C.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in droneyolo.h but not exported"
    assert declared <= set(_C.SYMBOLS), f"ctypes table misses {declared - set(_C.SYMBOLS)}"
    assert lib.dy_version() >= 100
    assert _C.lib().dy_nms_workspace_bytes(256, 10, 34000, 0) >= 256 * 34000 * 8


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the C structs: sizes must equal what gcc computes from the header."""
    src = '#include "droneyolo.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu\\n", sizeof(dy_conv_desc), sizeof(dy_decode_desc), sizeof(dy_nms_desc));return 0;}'
    exe = ROOT / "gpurun_out" / "_sizes"
    exe.parent.mkdir(exist_ok=True)
    subprocess.run(["gcc", "-x", "c", "-", "-I", str(ROOT / "include"), "-o", str(exe)], input=src, text=True, check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    from drone_yolo_b200 import _C

    assert sizes == [ctypes.sizeof(_C.ConvDesc), ctypes.sizeof(_C.DecodeDesc), ctypes.sizeof(_C.NmsDesc)]


def test_integration_md_struct_stubs_match_the_library_mirrors():
    """Every `class X(C.Structure)` stub printed in INTEGRATION.md (what a reference maintainer would paste) has the very
    fields, in order and of the same ctypes, of the mirror the package itself uses (a stale stub makes the library read
    past the struct)."""
    import re

    from drone_yolo_b200 import _C

    text = (ROOT / "INTEGRATION.md").read_text()
    stubs = re.findall(r"^class (\w+)\(C\.Structure\):.*?\n(    _fields_ = \[.*?\])[ \t]*(?:#[^\n]*)?$", text, flags=re.S | re.M)
    assert {n for n, _ in stubs} >= {"NmsDesc", "DecodeDesc"}
    for name, body in stubs:
        ns = {"C": ctypes}
        exec("class %s(C.Structure):\n%s\n" % (name, body), ns)
        mine = getattr(_C, name)
        assert ctypes.sizeof(ns[name]) == ctypes.sizeof(mine), name
        for (fa, ta), (fb, tb) in zip(ns[name]._fields_, mine._fields_):
            assert fa.rstrip("_") == fb.rstrip("_") and ctypes.sizeof(ta) == ctypes.sizeof(tb), (name, fa, fb)
        assert len(ns[name]._fields_) == len(mine._fields_), name


@pytest.mark.parametrize("cfg,n_params", sorted(PARAMS.items()))
def test_model_builds_with_reference_param_count(cfg, n_params):
    from drone_yolo_b200.nn.tasks import DetectionModel

    m = DetectionModel(cfg, nc=10, verbose=False)
    assert sum(p.numel() for p in m.parameters()) == n_params
    assert m.stride.tolist() == [4.0, 8.0, 16.0, 32.0]
    assert sorted(set(m.save)) == ([2, 4, 6, 9, 12, 15, 18, 21, 24, 27] if "sf" not in cfg else [0, 2, 4, 6, 9, 10, 13, 14, 17, 18, 21, 24, 27, 30])
    assert len(m.model) == (29 if "sf" not in cfg else 32)


def test_unknown_module_and_missing_yaml_fail():
    from drone_yolo_b200.nn.tasks import DetectionModel, parse_model

    with pytest.raises(FileNotFoundError):
        DetectionModel("yolov8n-nope.yaml", verbose=False)
    with pytest.raises(KeyError):
        parse_model({"nc": 3, "backbone": [[-1, 1, "GhostConv", [8, 3, 2]]], "head": []}, 3, verbose=False)


def test_repvgg_and_repconv_reparameterisation_algebra():
    """K3 + pad(K1) (+ identity) with BN folded equals the multi-branch forward (reference conv.py:206-247,
    block.py:1440-1478) — checked with the oracle's plain torch ops."""
    from drone_yolo_b200.nn.modules import RepConv, RepVGGBlock
    from oracle import recipe, torch_ref

    torch.manual_seed(3)
    for blk in (RepVGGBlock(16, 32, 3, 2), RepVGGBlock(16, 16, 3, 1), RepConv(16, 16, 3, 1, bn=True), RepConv(8, 24, 3, 2)):
        recipe.randomize_bn(blk, 5)
        blk.eval()
        x = torch.randn(2, blk.in_channels if hasattr(blk, "in_channels") else blk.c1, 12, 12)
        ref = torch_ref.module_forward(blk, x)
        w, b = blk.get_equivalent_kernel_bias()
        s = blk._geom()[1]
        got = torch.nn.functional.silu(torch.nn.functional.conv2d(x, w, b, stride=s, padding=1))
        torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-4)
        (blk.switch_to_deploy if isinstance(blk, RepVGGBlock) else blk.fuse_convs)()
        torch.testing.assert_close(torch_ref.module_forward(blk, x), ref, rtol=1e-4, atol=1e-4)


def test_conv_fuse_matches_bn_forward():
    from drone_yolo_b200.nn.modules import Conv
    from oracle import recipe, torch_ref

    torch.manual_seed(4)
    c = Conv(8, 16, 3, 2)
    recipe.randomize_bn(c, 9)
    c.eval()
    x = torch.randn(1, 8, 10, 10)
    ref = torch_ref.conv_forward(c, x)
    c.fuse()
    assert not hasattr(c, "bn")
    torch.testing.assert_close(torch_ref.conv_forward(c, x), ref, rtol=1e-4, atol=1e-5)


def test_pack_conv_weight_layout():
    from drone_yolo_b200 import kernels as K

    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    packed, bias = K.pack_conv_weight(w, torch.tensor([1.0, 2.0]))
    assert packed.shape == (9, 16, 64) and bias.shape == (16,)
    assert packed.dtype == torch.bfloat16
    # tap (r, s) row co, column ci  ==  w[co, ci, r, s]
    assert float(packed[1 * 3 + 2, 1, 2]) == float(w[1, 2, 1, 2].to(torch.bfloat16))
    assert float(packed[:, 2:].abs().sum()) == 0 and float(packed[:, :, 3:].abs().sum()) == 0
    assert bias[:2].tolist() == [1.0, 2.0] and float(bias[2:].abs().sum()) == 0


def test_box_helpers_match_reference_formulas():
    from drone_yolo_b200.utils import ops

    b = torch.tensor([[50.0, 60.0, 20.0, 10.0]])
    assert ops.xywh2xyxy(b).tolist() == [[40.0, 55.0, 60.0, 65.0]]
    boxes = torch.tensor([[100.0, 120.0, 300.0, 400.0], [-5.0, 10.0, 700.0, 500.0]])
    out = ops.scale_boxes((640, 640), boxes.clone(), (480, 640))      # letterboxed 480x640 image: pad (0, 80), gain 1
    assert out.tolist() == [[100.0, 40.0, 300.0, 320.0], [0.0, 0.0, 640.0, 420.0]]
    assert ops.make_divisible(33, 8) == 40


@pytest.mark.parametrize("h,w,imgsz,auto", [(1080, 1920, 640, True), (720, 1280, 640, False), (480, 640, 640, True),
                                            (375, 500, 640, False), (333, 517, 416, True)])
def test_letterbox_oracle_vs_cv2(h, w, imgsz, auto):
    """The numpy restatement of LetterBox + cv2's 8-bit INTER_LINEAR against the host path that calls cv2 (the reference's
    own arithmetic): bit-exact when shrinking, within one level on < 0.1 % of the pixels when enlarging; the geometry
    helper of the predictor agrees with the oracle's."""
    from drone_yolo_b200.engine.predictor import letterbox, letterbox_geometry
    from oracle import letterbox_np

    im = np.random.default_rng(h + w).integers(0, 256, (h, w, 3), dtype=np.uint8)
    ref = np.ascontiguousarray(letterbox(im, (imgsz, imgsz), auto=auto)[..., ::-1].transpose(2, 0, 1))
    got = letterbox_np.letterbox_chw_rgb(im, (imgsz, imgsz), auto=auto)
    assert got.shape == ref.shape
    assert letterbox_geometry((h, w), (imgsz, imgsz), auto=auto) == letterbox_np.geometry((h, w), (imgsz, imgsz), auto=auto)
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    shrinking = min(imgsz / h, imgsz / w) <= 1.0
    if shrinking:
        assert d.max() == 0
    else:
        assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_checkpoint_unpickler_allow_list_is_exact():
    """a prefix allow-list ("anything under torch / numpy / types") resolves code-executing globals.  Only exact
    (module, name) pairs, torch.nn.modules.* Module classes and ultralytics.* stand-ins resolve; dotted names never do."""
    import io
    import pickle

    import torch.nn as nn

    from drone_yolo_b200.nn.ckpt import _Unpickler

    u = _Unpickler(io.BytesIO(b""))
    for mod, name in (("numpy.testing._private.utils", "runstring"), ("types", "FunctionType"), ("types", "CodeType"),
                      ("torch.hub", "load"), ("torch.utils.cpp_extension", "load_inline"), ("os", "system"), ("builtins", "eval"),
                      ("builtins", "getattr"), ("torch", "load"), ("torch.serialization", "load"), ("numpy", "load"),
                      ("torch.nn.modules.module", "register_module_forward_hook"), ("torch", "Tensor.__reduce_ex__")):
        with pytest.raises(pickle.UnpicklingError):
            u.find_class(mod, name)
    assert u.find_class("torch.nn.modules.conv", "Conv2d") is nn.Conv2d
    assert u.find_class("collections", "OrderedDict").__name__ == "OrderedDict"
    assert u.find_class("torch._utils", "_rebuild_tensor_v2") is torch._utils._rebuild_tensor_v2
    assert u.find_class("torch", "float16") is torch.float16
    stub = u.find_class("ultralytics.nn.modules.block", "C2f")
    assert issubclass(stub, nn.Module) and stub.__module__ == "ultralytics.nn.modules.block"


def test_checkpoint_of_another_scale_fails_loudly(golden_dir, tmp_path):
    """`load` keeps intersecting keys only; a checkpoint whose tensors do not match the model built from its YAML
    must raise instead of predicting with random weights (the reference only logs 'Transferred x/y items')."""
    from drone_yolo_b200 import YOLO
    from drone_yolo_b200._C import DroneYoloError
    from drone_yolo_b200.nn.tasks import DetectionModel

    torch.manual_seed(0)
    m = DetectionModel("yolov8n-p2-repvgg.yaml", nc=10, verbose=False)
    sd = m.state_dict()
    ok = tmp_path / "ok.pt"
    torch.save({"yaml": "yolov8n-p2-repvgg.yaml", "model": sd, "nc": 10}, ok)
    assert YOLO(str(ok)).model.transferred[0] == len(sd)
    bad = tmp_path / "bad.pt"
    torch.save({"yaml": "yolov8s-p2-repvgg.yaml", "model": sd, "nc": 10}, bad)       # n-scale tensors, s-scale YAML
    with pytest.raises(DroneYoloError, match="tensors match"):
        YOLO(str(bad))
    fused = tmp_path / "fused.pt"
    torch.save({"yaml": "yolov8n-p2-repvgg.yaml", "model": {k: v for k, v in sd.items() if ".bn." not in k}, "nc": 10}, fused)
    with pytest.raises(DroneYoloError, match="fuse"):
        YOLO(str(fused))


def test_reference_pickled_checkpoint_ingestion(golden_dir):
    """A `.pt` written by the REAL reference (tools/make_ckpt_fixture.py: torch.save of its DetectionModel, the format
    mix6.py:18 loads) is read without the reference package: same YAML, names and weights (nn/tasks.py:786-926)."""
    import hashlib
    import json

    from drone_yolo_b200 import YOLO
    from drone_yolo_b200.nn.ckpt import load_reference_checkpoint

    meta = json.loads((golden_dir / "ref_ckpt_tiny.json").read_text())
    cfg, state, names, train_args = load_reference_checkpoint(str(golden_dir / "ref_ckpt_tiny.pt"))
    assert cfg["scale"] == "t" and cfg["nc"] == 10 and len(state) == meta["n_tensors"] and train_args["task"] == "detect"
    assert {int(k): v for k, v in names.items()} == {int(k): v for k, v in meta["names"].items()}

    def digest(sd):
        h = hashlib.sha256()
        for k in sorted(sd):
            h.update(k.encode())
            h.update(sd[k].detach().float().contiguous().numpy().tobytes())
        return h.hexdigest()

    assert digest(state) == meta["digest"]
    model = YOLO(str(golden_dir / "ref_ckpt_tiny.pt"))                       # the user-facing path (engine/model.py:266-302)
    assert digest(model.model.state_dict()) == meta["digest"]
    assert model.names[3] == "cls3"
    with pytest.raises(Exception):                                            # anything outside ultralytics / torch is refused
        import io
        import pickle

        class Evil:
            def __reduce__(self):
                return (os.system, ("true",))

        buf = io.BytesIO()
        pickle.dump({"model": Evil()}, buf)
        p = golden_dir.parent / "_evil_tmp.pt"
        p.write_bytes(buf.getvalue())
        try:
            load_reference_checkpoint(str(p))
        finally:
            p.unlink()


def test_cpu_tensors_fail_loudly():
    from drone_yolo_b200 import YOLO
    from drone_yolo_b200._C import DroneYoloError
    from drone_yolo_b200.nn.tasks import DetectionModel
    from drone_yolo_b200.utils import ops

    with pytest.raises(DroneYoloError):
        ops.non_max_suppression(torch.rand(1, 14, 100), 0.25, 0.45)
    with pytest.raises(AssertionError):
        ops.non_max_suppression(torch.rand(1, 14, 100), 1.5, 0.45)
    m = DetectionModel("yolov8n-p2-repvgg.yaml", nc=10, verbose=False).eval()
    with pytest.raises(DroneYoloError):
        m(torch.rand(1, 3, 64, 64))
    with pytest.raises(DroneYoloError):
        YOLO(m).predict(torch.rand(1, 3, 64, 64), device="cpu")


def test_shard_bounds_cover_the_batch():
    from drone_yolo_b200.parallel import shard_bounds

    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"])
from drone_yolo_b200.parallel import DetectionGather, shard_batch, split_detections
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['PORT']}", rank=int(os.environ["RANK"]), world_size=2)
rank = dist.get_rank()
batch = torch.arange(6 * 3, dtype=torch.float32).reshape(6, 3)
mine = shard_batch(batch, 2, rank)
assert mine.shape[0] == 3 and float(mine[0, 0]) == rank * 9
out = torch.full((3, 4, 6), float(rank + 1)); counts = torch.tensor([1 + rank, 2, 0], dtype=torch.int32)
g = DetectionGather(3, 4, "cpu")
oa, ca = g.gather(out, counts)
assert ca.tolist() == [1, 2, 0, 2, 2, 0], ca
dets = split_detections(oa, ca)
assert [d.shape[0] for d in dets] == [1, 2, 0, 2, 2, 0] and float(dets[3][0, 0]) == 2.0 and float(dets[0][0, 0]) == 1.0
dist.barrier(); dist.destroy_process_group(); print("ok", rank)
"""


def test_two_rank_gather_over_gloo(tmp_path):
    import socket

    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    script = tmp_path / "worker.py"
    script.write_text(GLOO_WORKER)
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), PORT=str(port), REPO=str(ROOT), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    for p in procs:
        out, _ = p.communicate(timeout=180)
        assert p.returncode == 0, out


def test_batched_postprocess_equals_per_image_scale_boxes():
    """DetectionPredictor.postprocess rescales the whole (B, max_det, 6) block at once; every row must equal the reference's
    per-image scale_boxes arithmetic (ops.py:92-127) bit for bit: tensor source, same-shape frames and ragged frames."""
    from types import SimpleNamespace

    from drone_yolo_b200.engine.predictor import DetectionPredictor
    from drone_yolo_b200.utils import ops

    p = DetectionPredictor(overrides=dict(max_det=50))
    p.model = SimpleNamespace(names={i: f"c{i}" for i in range(10)})
    B = 9
    torch.manual_seed(3)
    out = torch.rand(B, 50, 6) * 700 - 30
    counts = torch.tensor([0, 1, 50, 17, 3, 49, 50, 0, 25], dtype=torch.int32)
    img = torch.zeros(B, 3, 640, 640, dtype=torch.uint8)
    g = np.random.default_rng(0)
    ragged = [np.zeros((int(g.integers(100, 3000)), int(g.integers(100, 3000)), 3), np.uint8) for _ in range(B)]
    for orig in (img, [np.zeros((1080, 1920, 3), np.uint8)] * B, ragged):
        keep = out.clone()
        res = p.postprocess((out, counts), img, orig, [f"im{i}" for i in range(B)])
        assert torch.equal(out, keep)                                        # the caller's tensor is not modified
        for i, r in enumerate(res):
            osh = (640, 640) if isinstance(orig, torch.Tensor) else orig[i].shape[:2]
            rows = out[i, : int(counts[i])].clone()
            rows[:, :4] = ops.scale_boxes((640, 640), rows[:, :4], osh)
            assert torch.equal(rows, r.boxes.data) and r.orig_shape == tuple(osh) and len(r) == int(counts[i])
            if len(r):
                assert float(r.boxes.xyxy.min()) >= 0 and float(r.boxes.xyxy[:, [0, 2]].max()) <= osh[1]


def test_file_sources_follow_the_reference_loader(tmp_path):
    """Directory / glob / .txt / list-of-paths / video sources (LoadImagesAndVideos, data/loaders.py:284-446): sorted files, images
    before videos, batches of `batch`, an image batch never runs into a video, frames every `vid_stride`."""
    import cv2

    from drone_yolo_b200.engine.predictor import DetectionPredictor

    g = np.random.default_rng(0)
    for name, hw in (("b.png", (40, 60)), ("a.jpg", (50, 30)), ("c.bmp", (40, 60))):
        assert cv2.imwrite(str(tmp_path / name), g.integers(0, 256, (*hw, 3), dtype=np.uint8))
    (tmp_path / "notes.md").write_text("not an image")
    p = DetectionPredictor(overrides=dict(batch=2))
    got = list(p._batches(str(tmp_path)))
    assert [[Path(x).name for x in b[0]] for b in got] == [["a.jpg", "b.png"], ["c.bmp"]]
    assert got[0][1][0].shape == (50, 30, 3) and got[0][1][1].shape == (40, 60, 3) and got[0][2] is None
    assert np.array_equal(got[1][1][0], cv2.imread(str(tmp_path / "c.bmp")))
    assert [Path(x).name for b in p._batches(str(tmp_path / "*.png")) for x in b[0]] == ["b.png"]
    assert [Path(x).name for b in p._batches([str(tmp_path / "c.bmp"), tmp_path / "a.jpg"]) for x in b[0]] == ["a.jpg", "c.bmp"]
    (tmp_path / "list.txt").write_text("b.png\na.jpg\n")
    assert [Path(x).name for b in p._batches(str(tmp_path / "list.txt")) for x in b[0]] == ["a.jpg", "b.png"]
    with pytest.raises(FileNotFoundError):
        list(p._batches(str(tmp_path / "missing.jpg")))
    with pytest.raises(FileNotFoundError):
        sub = tmp_path / "empty"
        sub.mkdir()
        (sub / "x.md").write_text("")
        list(p._batches(str(sub)))
    with pytest.raises(TypeError):
        list(p._batches([str(tmp_path / "a.jpg"), np.zeros((8, 8, 3), np.uint8)]))
    # in-memory images stay ONE batch whatever `batch` says (LoadPilAndNumpy)
    arrs = [np.zeros((8, 8, 3), np.uint8)] * 5
    assert [len(b[1]) for b in p._batches(arrs)] == [5]

    vid = tmp_path / "clip.avi"
    wr = cv2.VideoWriter(str(vid), cv2.VideoWriter_fourcc(*"MJPG"), 10, (64, 48))
    if not wr.isOpened():
        pytest.skip("this OpenCV build cannot write MJPG/avi")
    for i in range(7):
        wr.write(np.full((48, 64, 3), 30 * i, np.uint8))
    wr.release()
    got = list(p._batches(str(vid)))
    assert [len(b[1]) for b in got] == [2, 2, 2, 1] and got[0][1][0].shape == (48, 64, 3)
    assert abs(int(got[3][1][0].mean()) - 180) <= 3                          # the last frame is the 7th
    got = list(p._batches(str(tmp_path)))                                    # images first, then the video; batches do not mix
    assert [len(b[1]) for b in got] == [2, 1, 2, 2, 2, 1] and Path(got[2][0][0]).name == "clip.avi"
    p3 = DetectionPredictor(overrides=dict(batch=4, vid_stride=3))
    got = list(p3._batches(str(vid)))
    assert [len(b[1]) for b in got] == [2] and abs(int(got[0][1][1].mean()) - 150) <= 3      # frames 3 and 6 (1-based)


def test_inference_slicer_argument_checks():
    """Constructor mirrors supervision.InferenceSlicer (mix6.py:84-89); bad arguments fail before any CUDA work."""
    from drone_yolo_b200 import InferenceSlicer
    from drone_yolo_b200._C import DroneYoloError
    from drone_yolo_b200.engine.slicer import generate_offsets

    s = InferenceSlicer(None, slice_wh=(2160, 2160), overlap_ratio_wh=(0.2, 0.2), iou_threshold=0.7, thread_workers=1, conf=0.7, classes=[0])
    assert s.slice_wh == (2160, 2160) and s.iou_threshold == 0.7 and s.predict_kwargs == {"conf": 0.7, "classes": [0]}
    with pytest.raises(AssertionError):
        InferenceSlicer(None, iou_threshold=1.5)
    for bad in (np.zeros((8, 8), np.uint8), np.zeros((8, 8, 3), np.float32), "frame.jpg"):
        with pytest.raises(DroneYoloError):
            s(bad)
    with pytest.raises(ValueError):
        generate_offsets((100, 100), (0, 10), (0.2, 0.2))
    with pytest.raises(ValueError):
        generate_offsets((100, 100), (10, 10), (1.0, 0.2))
    assert generate_offsets((100, 50), (200, 200), (0.2, 0.2)).tolist() == [[0, 0, 100, 50]]      # frame smaller than a tile


def test_frame_stream_batching_fast_and_ragged_paths():
    """DetectionPredictor._iter_batches: a stream of equally shaped frames is cut into `batch`-sized lists (one islice per batch);
    a change of frame shape closes the batch early; a 4-D tensor or a list inside the stream is its own batch; names count on."""
    from types import SimpleNamespace

    from drone_yolo_b200.engine.predictor import DetectionPredictor

    p = DetectionPredictor.__new__(DetectionPredictor)
    p.args = SimpleNamespace(batch=4)
    a, b = np.zeros((4, 6, 3), np.uint8), np.zeros((8, 6, 3), np.uint8)
    got = [(len(i), i[0].shape, pa[0], pa[-1]) for pa, i, _ in p._iter_batches(iter([a] * 9))]
    assert got == [(4, a.shape, "image0.jpg", "image3.jpg"), (4, a.shape, "image4.jpg", "image7.jpg"), (1, a.shape, "image8.jpg", "image8.jpg")]
    got = [(len(i), i[0].shape, pa[0]) for pa, i, _ in p._iter_batches(iter([a] * 5 + [b] * 3 + [a] * 2))]
    assert got == [(4, a.shape, "image0.jpg"), (1, a.shape, "image4.jpg"), (3, b.shape, "image5.jpg"), (2, a.shape, "image8.jpg")]
    got = [(len(i), pa[0]) for pa, i, _ in p._iter_batches(iter([a, a, [b, b, b], a]))]
    assert got == [(2, "image0.jpg"), (3, "image2.jpg"), (1, "image5.jpg")]
    assert list(p._iter_batches(iter([]))) == []


def test_results_are_lazy_views_of_one_block():
    """Results built from (block, image, count): len() without touching the rows, boxes sliced on first access, `_with` /
    the boxes setter keep working, and the rows are a copy (the predictor's pinned read-back buffer is reused)."""
    from drone_yolo_b200.engine.results import Boxes, Results

    block = torch.arange(2 * 5 * 6, dtype=torch.float32).reshape(2, 5, 6)
    r = Results(np.zeros((8, 8, 3), np.uint8), "image1.jpg", {0: "a"}, boxes=(block, 1, 3), orig_shape=(8, 8))
    assert len(r) == 3 and r._boxes is None
    assert torch.equal(r.boxes.data, block[1, :3]) and r.boxes.conf.shape == (3,) and r.boxes.xywh.shape == (3, 4)
    assert len(r.numpy()) == 3 and isinstance(r.numpy().boxes.data, np.ndarray)
    r.boxes = Boxes(block[0, :2], (8, 8))
    assert len(r) == 2
    assert len(Results(np.zeros((8, 8, 3), np.uint8), "x", {}, boxes=None)) == 0
    assert len(Results(np.zeros((8, 8, 3), np.uint8), "x", {}, boxes=torch.zeros(0, 6))) == 0


@pytest.mark.parametrize("yaml_name,expect", [("yolov8n-p2-repvgg.yaml", 3), ("yolov8s-p2-repvgg.yaml", 3), ("yolov8n-p2-repvgg-sf.yaml", 3)])
def test_plan_splits_cv1_across_the_upsample_symbolically(yaml_name, expect):
    """The plan's symbolic pass (no GPU): in the Drone-YOLO graphs every `Upsample -> Concat -> C2f` of the top-down neck is lowered as
    a low-resolution 1x1 conv (fp32, no activation) + a 1x1 conv over the skip tensor with `pre`; no op carries a fused upsample store."""
    from drone_yolo_b200.engine.plan import LayerPlan
    from drone_yolo_b200.nn.tasks import DetectionModel

    torch.manual_seed(0)
    model = DetectionModel(yaml_name, nc=10, verbose=False).eval().fuse(verbose=False)
    lp = LayerPlan.__new__(LayerPlan)
    lp.model, lp.mb, lp.H, lp.W, lp.device = model, 2, 128, 128, None
    lp.bufs, lp.ops, lp.keep = [], [], []
    lp.fuse_upsample = lp.fuse_tail = lp.fuse_decode = lp.fuse_cv1 = True
    lp.split_up, lp.up_split = True, {}
    lp.head_lanes, lp.lane = 0, 0
    lp._build_symbolic()
    pre = [op for op in lp.ops if op.get("pre") is not None]
    assert len(pre) == expect == len(lp.up_split) and not any(op.get("up") is not None for op in lp.ops)
    for op in pre:
        t = op["pre"]
        assert lp.bufs[t.buf].esz == 4 and (2 * t.H, 2 * t.W) == (op["out"].H, op["out"].W) and op["k"] == 1 and op["act"]
        prod = [o for o in lp.ops if o.get("out") == t]
        assert len(prod) == 1 and prod[0]["k"] == 1 and not prod[0]["act"]
    names = lp.describe()
    assert sum("+pre(up)" in nm for nm in names) == expect and len(names) == len(lp.ops)
    # with the split switched off the same graph carries the fused upsample stores instead
    lp2 = LayerPlan.__new__(LayerPlan)
    lp2.__dict__.update({k: v for k, v in lp.__dict__.items() if k not in ("bufs", "ops", "keep", "up_split")})
    lp2.bufs, lp2.ops, lp2.keep, lp2.split_up, lp2.up_split = [], [], [], False, {}
    lp2._build_symbolic()
    assert sum(op.get("up") is not None for op in lp2.ops) == expect and not any(op.get("pre") is not None for op in lp2.ops)


def test_rescale_params_match_scale_boxes_arithmetic():
    """ops.rescale_params (the per-image block the NMS output phase reads, dy_nms_desc.rescale): pad / gain / clamp values are
    the ones `scale_boxes` + `clip_boxes` use (reference ops.py:92-127, 335-354), for mixed shapes incl. repeated ones."""
    from drone_yolo_b200.utils import ops

    shapes = [(1080, 1920), (480, 640), (1080, 1920), (333, 777), (640, 640)]
    rs = ops.rescale_params((640, 640), shapes)
    assert rs.shape == (5, 8) and rs.dtype == torch.float32 and torch.equal(rs[0], rs[2])
    for i, s in enumerate(shapes):
        gain = min(640 / s[0], 640 / s[1])
        pad = (round((640 - s[1] * gain) / 2 - 0.1), round((640 - s[0] * gain) / 2 - 0.1))
        assert rs[i, :5].tolist() == [float(pad[0]), float(pad[1]), float(np.float32(gain)), float(s[1]), float(s[0])]
        box = torch.tensor([[100.0, 120.0, 500.0, 630.0]])
        ref = ops.scale_boxes((640, 640), box.clone(), s)
        got = box.clone()
        got[:, [0, 2]] -= rs[i, 0]
        got[:, [1, 3]] -= rs[i, 1]
        got /= rs[i, 2]
        got[:, [0, 2]] = got[:, [0, 2]].clamp(0, float(rs[i, 3]))
        got[:, [1, 3]] = got[:, [1, 3]].clamp(0, float(rs[i, 4]))
        assert torch.equal(got, ref)
