"""CPU tests: the oracle (oracle/) against the golden vectors produced by the real reference (tools/make_golden.py),
and the host-side model construction against the reference's state_dict digest."""
import numpy as np
import pytest
import torch

from oracle import decode_np, nms_np, recipe, torch_ref

NMS_CASES = {
    "default": dict(conf_thres=0.001, iou_thres=0.7, max_det=300),
    "multilabel": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),
    "agnostic": dict(conf_thres=0.001, iou_thres=0.5, max_det=100, agnostic=True),
    "classes": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, classes=[1, 3, 7]),
    "maxnms": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=200),
    "predict": dict(conf_thres=None, iou_thres=0.45, max_det=300),       # conf: the fixture's `predict_conf` (a few dozen rows survive)
}
REGIMES = ("sparse", "vallike", "dense")
CONV_FIXTURES = ("n_repvgg_128", "n_repvgg_sf_64", "n_p2_64", "s_repvgg_64")


def nms_case(g, case):
    kw = dict(NMS_CASES[case])
    if kw["conf_thres"] is None:
        kw["conf_thres"] = float(g["predict_conf"])
    return kw


def frames_of(g):
    """The raw uint8 BGR frames of a predict fixture, regenerated from its seeds."""
    return recipe.predict_frames([tuple(v) for v in g["shapes"].tolist()], int(g["frame_seed0"]))


def build_model(g):
    from drone_yolo_b200.nn.tasks import DetectionModel

    torch.manual_seed(int(g["model_seed"]))
    m = DetectionModel(str(g["yaml"]), nc=int(g["nc"]), verbose=False)
    recipe.apply_recipe(m, int(g["bn_seed"]), float(g["cls_delta"]), float(g["cls_gain"]) if "cls_gain" in g.files else 1.0)
    return m.eval()


@pytest.mark.parametrize("regime", REGIMES)
def test_decode_oracle_matches_reference(golden_dir, regime):
    g = np.load(golden_dir / f"decode_nms_{regime}.npz")
    raw = recipe.synthetic_raw_maps(int(g["B"]), int(g["imgsz"]), int(g["nc"]), float(g["mu"]), int(g["raw_seed"]))
    y = decode_np.decode([r.numpy() for r in raw], [4.0, 8.0, 16.0, 32.0], int(g["nc"]))
    assert y.shape == g["y"].shape
    # boxes in pixels, probabilities in [0,1]: the numpy restatement tracks the reference's aten ops to float rounding
    np.testing.assert_allclose(y[:, :4], g["y"][:, :4], rtol=0, atol=2e-4)
    np.testing.assert_allclose(y[:, 4:], g["y"][:, 4:], rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("regime", REGIMES)
@pytest.mark.parametrize("case", sorted(NMS_CASES))
def test_nms_oracle_bit_exact_vs_reference(golden_dir, regime, case):
    g = np.load(golden_dir / f"decode_nms_{regime}.npz")
    out, kept = nms_np.non_max_suppression(g["y"], return_kept=True, **nms_case(g, case))
    for b in range(int(g["B"])):
        ref_rows, ref_kept = g[f"{case}_out{b}"], g[f"{case}_kept{b}"]
        assert out[b].shape == ref_rows.shape, (regime, case, b)
        assert np.array_equal(out[b].view(np.uint32), ref_rows.view(np.uint32)), (regime, case, b)   # bit-exact rows
        assert np.array_equal(kept[b], ref_kept), (regime, case, b)                                   # and kept indices
    if case == "predict":
        assert min(g[f"predict_out{b}"].shape[0] for b in range(int(g["B"]))) >= 20, "the predict case must not be vacuous"


@pytest.mark.parametrize("regime", REGIMES)
@pytest.mark.parametrize("case", ["default", "multilabel", "predict"])
def test_nms_oracle_bit_exact_vs_reference_34k(golden_dir, regime, case):
    """BASELINE config 4 at its real size (34 000 anchors, one image), rows and kept indices written by the real reference.
    Dense: 33 5xx candidates > max_nms, so the reference's truncation ran (tools/make_golden.py asserted that its unstable
    argsort gave the stable order's result on this vector)."""
    g = np.load(golden_dir / f"decode_nms_{regime}_34k.npz")
    assert g["y"].shape == (1, 14, 34000)
    out, kept = nms_np.non_max_suppression(g["y"], return_kept=True, **nms_case(g, case))
    assert np.array_equal(out[0].view(np.uint32), g[f"{case}_out0"].view(np.uint32)), (regime, case)
    assert np.array_equal(kept[0], g[f"{case}_kept0"]), (regime, case)
    if regime == "dense" and case == "default":
        assert int((g["y"][0, 4:].max(0) > 0.001).sum()) > 30000


@pytest.mark.parametrize("tag", ["ragged", "rect"])
def test_predict_pre_and_post_processing_vs_reference(golden_dir, tag):
    """The reference's YOLO.predict on raw frames, with the conv stack factored out: (1) the letterbox restatement reproduces
    the tensor its preprocess handed to the model, (2) the NMS oracle on the tensor its model handed to NMS followed by THIS
    package's scale_boxes / clip_boxes (the host post-step the predictor runs) reproduces its `boxes.data`, bit for bit."""
    from oracle import letterbox_np
    from drone_yolo_b200.utils import ops

    g = np.load(golden_dir / f"predict_{tag}.npz")
    frames = frames_of(g)
    same = len({f.shape for f in frames}) == 1
    imgsz = int(g["imgsz"])
    canv = np.stack([letterbox_np.letterbox_chw_rgb(f, (imgsz, imgsz), auto=same) for f in frames])
    assert canv.shape == g["im_u8"].shape and np.array_equal(canv, g["im_u8"])
    out = nms_np.non_max_suppression(g["y"], float(g["conf"]), float(g["iou"]), max_det=int(g["max_det"]))
    for b, rows in enumerate(out):
        t = torch.from_numpy(rows.copy())
        t[:, :4] = ops.scale_boxes(canv.shape[2:], t[:, :4], tuple(int(v) for v in g[f"orig_shape{b}"]))
        want = g[f"boxes{b}"]
        assert t.shape == want.shape and want.shape[0] >= 10
        assert np.array_equal(t.numpy().view(np.uint32), want.view(np.uint32)), (tag, b)
        padded = torch.zeros((1, int(g["max_det"]), 6))                      # the predictor's batched form of the same arithmetic
        padded[0, : rows.shape[0]] = torch.from_numpy(rows)
        ops.scale_boxes_batch(canv.shape[2:], padded[..., :4], [tuple(int(v) for v in g[f"orig_shape{b}"])])
        assert np.array_equal(padded[0, : rows.shape[0]].numpy().view(np.uint32), want.view(np.uint32)), (tag, b)


def test_convstack_oracle_matches_reference_at_640(golden_dir):
    """BASELINE config 2's model (Drone-YOLO-s) at 640x640, one image: raw maps stored as fp16."""
    g = np.load(golden_dir / "convstack_s_repvgg_640_big.npz")
    m = build_model(g)
    x = recipe.images(1, 640, 640, int(g["image_seed"]))
    y, raw = torch_ref.forward(m, x)
    assert y.shape == (1, 14, 34000)
    for i, r in enumerate(raw):
        np.testing.assert_allclose(r.numpy(), g[f"raw{i}"].astype(np.float32), rtol=2e-3, atol=2e-3)      # fp16 storage
    np.testing.assert_allclose(y[:, :4], g["y"][:, :4], rtol=0, atol=1e-2)
    np.testing.assert_allclose(y[:, 4:], g["y"][:, 4:], rtol=1e-3, atol=1e-6)


LABELS = [[[3.0, 40.0, 52.0, 30.0, 22.0], [7.0, 90.5, 30.25, 12.0, 44.0], [3.0, 41.0, 51.0, 28.0, 24.0]], []]   # tools/make_golden.py


@pytest.mark.parametrize("regime", REGIMES)
def test_nms_oracle_apriori_labels_vs_reference(golden_dir, regime):
    """Validator form (multi_label) with a-priori labels appended to the candidates (ops.py:272-277)."""
    g = np.load(golden_dir / f"decode_nms_{regime}.npz")
    out, kept = nms_np.non_max_suppression(g["y"], 0.001, 0.7, multi_label=True, max_det=300, labels=LABELS, return_kept=True)
    for b in range(int(g["B"])):
        assert np.array_equal(out[b].view(np.uint32), g[f"labels_out{b}"].view(np.uint32)), (regime, b)
        assert np.array_equal(kept[b], g[f"labels_kept{b}"]), (regime, b)
    if regime != "dense":                       # the two score-1.0 labels of image 0 that survive NMS lead its rows
        assert out[0][0, 4] == 1.0 and out[0][0, 5] in (3.0, 7.0)


def test_nms_restatement_equals_installed_torchvision(golden_dir):
    import torchvision

    g = np.load(golden_dir / "decode_nms_vallike.npz")

    def tv(boxes, scores, thr):
        return torchvision.ops.nms(torch.from_numpy(boxes), torch.from_numpy(scores), thr).numpy()

    a = nms_np.non_max_suppression(g["y"], conf_thres=0.001, iou_thres=0.7, return_kept=True)
    b = nms_np.non_max_suppression(g["y"], conf_thres=0.001, iou_thres=0.7, return_kept=True, nms_fn=tv)
    for x, y in zip(a[1], b[1]):
        assert np.array_equal(x, y)


def test_nms_tie_and_threshold_rules():
    # identical boxes, tied scores: the lower index survives; IoU == thr is NOT suppressed (strict >)
    boxes = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [0, 0, 10, 7]], np.float32)
    scores = np.array([0.5, 0.5, 0.4], np.float32)
    assert nms_np.nms_greedy(boxes, scores, 0.7).tolist() == [0, 2]      # IoU(0,2) = 0.7 exactly -> kept
    assert nms_np.nms_greedy(boxes, scores, 0.69).tolist() == [0]
    assert nms_np.nms_greedy(np.zeros((0, 4), np.float32), np.zeros((0,), np.float32), 0.5).size == 0


@pytest.mark.parametrize("tag", CONV_FIXTURES)
def test_seeded_construction_matches_reference_state(golden_dir, tag):
    g = np.load(golden_dir / f"convstack_{tag}.npz")
    m = build_model(g)
    assert len(m.state_dict()) == int(g["n_keys"])
    assert sum(p.numel() for p in m.parameters()) == int(g["n_params"])
    assert np.array_equal(m.stride.numpy(), g["stride"])
    assert recipe.state_digest(m) == str(g["state_digest"])


@pytest.mark.parametrize("tag", CONV_FIXTURES)
def test_convstack_oracle_matches_reference(golden_dir, tag):
    g = np.load(golden_dir / f"convstack_{tag}.npz")
    m = build_model(g)
    x = recipe.images(int(g["B"]), int(g["imgsz"]), int(g["imgsz"]), int(g["image_seed"]))
    y, raw = torch_ref.forward(m, x)
    for i, r in enumerate(raw):
        np.testing.assert_allclose(r.numpy(), g[f"raw{i}"].astype(np.float32), rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(y[:, :4], g["y"][:, :4], rtol=0, atol=1e-2)
    np.testing.assert_allclose(y[:, 4:], g["y"][:, 4:], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("tag", ("n_repvgg_128", "n_repvgg_sf_64"))
def test_fuse_reparameterisation_is_exact_enough(golden_dir, tag):
    """fuse() (BN fold + RepVGG merge) must not change the oracle's outputs beyond fp32 noise."""
    g = np.load(golden_dir / f"convstack_{tag}.npz")
    m = build_model(g)
    x = recipe.images(int(g["B"]), int(g["imgsz"]), int(g["imgsz"]), int(g["image_seed"]))
    y0, raw0 = torch_ref.forward(m, x)
    m.fuse(verbose=False)
    assert m.is_fused() and not any(isinstance(k, torch.nn.BatchNorm2d) for k in m.modules())
    y1, raw1 = torch_ref.forward(m, x)
    for a, b in zip(raw0, raw1):
        np.testing.assert_allclose(a.numpy(), b.numpy(), rtol=1e-3, atol=2e-3)
    np.testing.assert_allclose(y0[:, :4], y1[:, :4], atol=2e-2)


# ---------------------------------------------------------------------------------------------- tiled frames (mix6.py:84-89)
def test_slicer_offsets_known_answers():
    """Hand-computed tile grids: the call site's own arguments on a 4K frame (mix6.py:84-87), and an exact-fit frame."""
    from oracle import slicer_np
    from drone_yolo_b200.engine.slicer import generate_offsets

    want = [[0, 0, 2160, 2160], [1728, 0, 3840, 2160], [3456, 0, 3840, 2160],
            [0, 1728, 2160, 2160], [1728, 1728, 3840, 2160], [3456, 1728, 3840, 2160]]      # stride 2160 - int(0.2 * 2160) = 1728
    assert slicer_np.generate_offsets((3840, 2160), (2160, 2160), (0.2, 0.2)).tolist() == want
    assert generate_offsets((3840, 2160), (2160, 2160), (0.2, 0.2)).tolist() == want
    assert slicer_np.generate_offsets((640, 640), (640, 640), (0.0, 0.0)).tolist() == [[0, 0, 640, 640]]
    g = np.random.default_rng(0)
    for _ in range(50):
        wh = (int(g.integers(1, 5000)), int(g.integers(1, 5000)))
        sl = (int(g.integers(1, 3000)), int(g.integers(1, 3000)))
        ov = (float(g.uniform(0, 0.9)), float(g.uniform(0, 0.9)))
        a, b = slicer_np.generate_offsets(wh, sl, ov), generate_offsets(wh, sl, ov)
        assert np.array_equal(a, b) and a.dtype == b.dtype == np.int64
        assert a[:, 2].max() == wh[0] and a[:, 3].max() == wh[1] and (a[:, 2] > a[:, 0]).all() and (a[:, 3] > a[:, 1]).all()


def test_slicer_merge_nms_known_answers():
    from oracle import slicer_np

    box = [10.0, 10.0, 50.0, 50.0]
    rows = np.array([box + [0.5, 0], box + [0.9, 0], box + [0.8, 1], [200, 200, 240, 240, 0.1, 0]], dtype=np.float64)
    assert slicer_np.box_nms_keep(rows, 0.5).tolist() == [False, True, True, True]          # same box: best of each category
    assert slicer_np.box_nms_keep(rows, 0.5, class_agnostic=True).tolist() == [False, True, False, True]
    tie = np.array([box + [0.7, 0], box + [0.7, 0]], dtype=np.float64)
    assert slicer_np.box_nms_keep(tie, 0.5).tolist() == [False, True]                       # equal conf: the higher row ranks first
    half = np.array([[0, 0, 10, 10, 0.9, 0], [0, 0, 10, 5, 0.8, 0]], dtype=np.float64)      # IoU exactly 0.5
    assert slicer_np.box_nms_keep(half, 0.5).tolist() == [True, True]                       # strict >
    assert slicer_np.box_nms_keep(half, 0.4999).tolist() == [True, False]
    flat = np.array([[5, 5, 5, 5, 0.9, 0], [5, 5, 5, 5, 0.8, 0]], dtype=np.float64)         # zero area: IoU is NaN, nothing removed
    assert slicer_np.box_nms_keep(flat, 0.5).tolist() == [True, True]
    assert slicer_np.box_nms_keep(np.zeros((0, 6)), 0.5).shape == (0,)
    # an object in the overlap of two tiles is found twice; after the move to frame coordinates the weaker copy goes
    t0 = np.array([[1800, 100, 1900, 200, 0.8, 0]], dtype=np.float32)
    t1 = np.array([[72.5, 100, 172, 200, 0.9, 0], [500, 500, 600, 600, 0.3, 2]], dtype=np.float32)
    offs = slicer_np.generate_offsets((3840, 2160), (2160, 2160), (0.2, 0.2))
    out = slicer_np.merge_tiles([t0, t1], offs[:2], 0.7)
    assert out.dtype == np.float64 and out.tolist() == [[1800.5, 100, 1900, 200, np.float32(0.9), 0], [2228, 500, 2328, 600, np.float32(0.3), 2]]


def test_slicer_merge_nms_agrees_with_torchvision_on_tie_free_rows():
    """Independent check of the greedy rule: on tie-free scores, the kept set per category equals torchvision.ops.nms on the same
    float64 boxes (same IoU formula, strict >); the two differ only in their tie order, which the known-answer test pins."""
    import torchvision
    from oracle import slicer_np

    g = np.random.default_rng(21)
    n = 600
    c = g.uniform(0, 800, (n, 2))
    wh = g.uniform(30, 120, (n, 2))
    conf = g.permutation(n).astype(np.float64) / n + 0.001                   # all distinct
    rows = np.concatenate([c - wh / 2, c + wh / 2, conf[:, None], g.integers(0, 3, (n, 1)).astype(np.float64)], 1)
    for thr in (0.3, 0.5, 0.7):
        keep = slicer_np.box_nms_keep(rows, thr)
        want = np.zeros(n, dtype=bool)
        for k in range(3):
            idx = np.nonzero(rows[:, 5] == k)[0]
            kept = torchvision.ops.nms(torch.from_numpy(rows[idx, :4]), torch.from_numpy(rows[idx, 4]), thr).numpy()
            want[idx[kept]] = True
        assert np.array_equal(keep, want) and 0 < keep.sum() < n
        t = torchvision.ops.nms(torch.from_numpy(rows[:, :4]), torch.from_numpy(rows[:, 4]), thr).numpy()
        agn = np.zeros(n, dtype=bool)
        agn[t] = True
        assert np.array_equal(slicer_np.box_nms_keep(rows, thr, class_agnostic=True), agn)


# ---------------------------------------------------------------------------------------------- validator caller
def _val_batch(g, device="cpu"):
    rp = g["ratio_pad"]
    B = int(g["B"])
    return {"img": torch.from_numpy(g["img"]).to(device), "cls": torch.from_numpy(g["cls"]).to(device),
            "bboxes": torch.from_numpy(g["bboxes"]).to(device), "batch_idx": torch.from_numpy(g["batch_idx"]).to(device),
            "ori_shape": [tuple(int(v) for v in g["ori_shape"])] * B,
            "ratio_pad": [((float(rp[0]), float(rp[1])), (float(rp[2]), float(rp[3])))] * B}


@pytest.mark.parametrize("tag", ["n128", "n128_hybrid"])
def test_validator_metrics_restatement_vs_reference(golden_dir, tag):
    """utils/metrics.py + the validator's update_metrics / get_stats against the REAL reference's DetectionValidator
    (tools/make_golden_val.py): from the reference's own NMS rows the correct matrix is identical and P / R / AP per class /
    mAP50 / mAP50-95 agree to 1e-12; from the stored stats `ap_per_class` alone agrees too."""
    from types import SimpleNamespace

    from drone_yolo_b200.engine.validator import DetectionValidator
    from drone_yolo_b200.utils.metrics import ap_per_class

    g = np.load(golden_dir / f"val_{tag}.npz")
    p, r, f1, ap, classes = ap_per_class(g["tp"], g["conf"], g["pred_cls"], g["target_cls"])
    np.testing.assert_allclose(ap, g["all_ap"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(p, g["p"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(r, g["r"], rtol=0, atol=1e-12)
    assert classes.tolist() == g["ap_class_index"].tolist()

    v = DetectionValidator(args=dict(conf=0.001, iou=0.7, max_det=300, save_hybrid=bool(g["save_hybrid"])))
    v.device = torch.device("cpu")
    v.init_metrics(SimpleNamespace(names={i: f"cls{i}" for i in range(int(g["nc"]))}))
    rows = np.split(g["rows"], np.cumsum(g["n_rows"])[:-1])
    batch = _val_batch(g)
    v.update_metrics([torch.from_numpy(x.copy()) for x in rows], batch)
    res = v.get_stats()
    assert np.array_equal(v.last_stats["tp"], g["tp"])
    assert np.array_equal(v.last_stats["conf"], g["conf"]) and np.array_equal(v.last_stats["pred_cls"], g["pred_cls"])
    assert np.array_equal(v.last_stats["target_cls"], g["target_cls"])
    got = np.array([res[k] for k in ("metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)", "fitness")])
    np.testing.assert_allclose(got, g["results"], rtol=0, atol=1e-12)
    assert 0.15 < g["results"][3] < 0.99                                  # the fixture is not vacuous


def test_validator_argument_checks():
    from drone_yolo_b200._C import DroneYoloError
    from drone_yolo_b200.engine.validator import DetectionValidator

    assert DetectionValidator().args.conf == 0.001                         # engine/validator.py:101-102
    for bad in (dict(half=True), dict(plots=True), dict(save_json=True)):
        with pytest.raises(DroneYoloError):
            DetectionValidator(args=bad)
    with pytest.raises(DroneYoloError):
        DetectionValidator()(model=None, batches=[])


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(12))
def test_metrics_restatement_vs_live_reference(seed):
    """utils/metrics.py against the LIVE reference functions (authoring container only; skipped where /root/reference is absent):
    box_iou, match_predictions (validator.py:224-264) and ap_per_class (metrics.py:537-623) on random detections / labels incl.
    empty sides, duplicate scores and classes without predictions."""
    from oracle import ref_shim

    ref_shim.load()
    from ultralytics.engine.validator import BaseValidator
    from ultralytics.utils import metrics as rm

    from drone_yolo_b200.utils import metrics as m

    g = torch.Generator().manual_seed(seed)
    n, l, nc = (0, 5, 4) if seed == 0 else ((7, 0, 4) if seed == 1 else (int(torch.randint(1, 60, (1,), generator=g)), int(torch.randint(1, 25, (1,), generator=g)), 6))

    def boxes(k):
        xy = torch.rand(k, 2, generator=g) * 80
        wh = torch.rand(k, 2, generator=g) * 40 + 2
        return torch.cat((xy, xy + wh), 1)

    gt = boxes(l)
    gt_cls = torch.randint(0, nc, (l,), generator=g).float()
    det = torch.cat((gt[torch.randint(0, max(l, 1), (n,), generator=g)] + torch.randn(n, 4, generator=g) * 3 if l else boxes(n), boxes(n)))[:n]
    pred_cls = torch.randint(0, nc, (n,), generator=g).float()
    conf = (torch.rand(n, generator=g) * 8).round() / 8                                   # duplicate scores on purpose
    iou_ref, iou_got = rm.box_iou(gt, det), m.box_iou(gt, det)
    assert torch.equal(iou_ref, iou_got)
    iouv = torch.linspace(0.5, 0.95, 10)
    v = BaseValidator.__new__(BaseValidator)
    v.iouv = iouv
    tp_ref = v.match_predictions(pred_cls, gt_cls, iou_ref)
    tp_got = m.match_predictions(pred_cls, gt_cls, iou_got, iouv)
    assert torch.equal(tp_ref, tp_got)
    if n and l:
        ref = rm.ap_per_class(tp_ref.numpy(), conf.numpy(), pred_cls.numpy(), gt_cls.numpy())
        p, r, f1, ap, classes = m.ap_per_class(tp_got.numpy(), conf.numpy(), pred_cls.numpy(), gt_cls.numpy())
        np.testing.assert_allclose(p, ref[2], rtol=0, atol=1e-12)
        np.testing.assert_allclose(r, ref[3], rtol=0, atol=1e-12)
        np.testing.assert_allclose(f1, ref[4], rtol=0, atol=1e-12)
        np.testing.assert_allclose(ap, ref[5], rtol=0, atol=1e-12)
        assert classes.tolist() == ref[6].tolist()


@pytest.mark.reference
@pytest.mark.parametrize("seed", range(8))
def test_nms_and_decode_oracles_vs_live_reference(seed):
    """The numpy oracles against the LIVE reference on fresh random inputs (authoring container only): Detect decode
    (head.py:100-131) to 1e-5 relative, ops.non_max_suppression (ops.py:181-332) rows bit for bit, for a random option set per seed
    (multi_label, agnostic, class filter, small max_det / max_nms, a-priori labels) and clustered boxes that really suppress."""
    from oracle import ref_shim

    tasks = ref_shim.load()
    from ultralytics.utils import ops

    g = torch.Generator().manual_seed(100 + seed)
    nc, B, imgsz = 10, 2, 64
    shapes = [(imgsz // s, imgsz // s) for s in (4, 8, 16, 32)]
    raw = [torch.cat((1.5 * torch.randn(B, 64, h, w, generator=g), -2.0 + 2.0 * torch.randn(B, nc, h, w, generator=g)), 1) for h, w in shapes]
    det = tasks.Detect(nc=nc, ch=(16, 32, 64, 128)).eval()
    det.stride = torch.tensor([4.0, 8.0, 16.0, 32.0])
    with torch.no_grad():
        y_ref = det._inference([r.clone() for r in raw])
    y = decode_np.decode([r.numpy() for r in raw], [4.0, 8.0, 16.0, 32.0], nc)
    np.testing.assert_allclose(y, y_ref.numpy(), rtol=1e-5, atol=1e-4)
    kw = dict(conf_thres=[0.001, 0.05, 0.25][seed % 3], iou_thres=[0.7, 0.45, 0.6, 0.3][seed % 4], max_det=[300, 20][seed % 2])
    if seed % 2:
        kw["multi_label"] = True
    if seed % 3 == 0:
        kw["agnostic"] = True
    if seed % 4 == 1:
        kw["classes"] = [1, 3, 7]
    if seed == 5:
        kw["max_nms"] = 50
    labels = [[[3.0, 20.0, 22.0, 10.0, 12.0]], []] if seed in (2, 7) else ()
    yr = y_ref.clone()
    ref = ops.non_max_suppression(yr, labels=[torch.tensor(l).reshape(-1, 5) for l in labels] if labels else (), **kw)
    got = nms_np.non_max_suppression(y_ref.numpy(), labels=labels, **kw)
    for a, b in zip(got, ref):
        assert a.shape == tuple(b.shape) and np.array_equal(a.view(np.uint32), b.numpy().view(np.uint32))
    assert sum(len(a) for a in got) > 0
