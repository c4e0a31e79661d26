"""GPU diagnostics: run each kernel family against its reference and print error statistics (no early exit).

    python tools/gpu_diag.py umma|conv|decode|nms|aux|model|all

Each family is meant to be run in its own process (a device-side trap poisons the CUDA context).
"""
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from drone_yolo_b200 import kernels as K  # noqa: E402
from oracle import decode_np, nms_np, recipe, torch_ref  # noqa: E402

dev = torch.device("cuda:0")


def report(name, got, ref, tol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    rel = err / (ref.abs() + 1e-3)
    bad = int((err > tol * (1 + ref.abs())).sum())
    print(f"  {name:<58} max_abs {err.max().item():.4e} max_rel {rel.max().item():.3e} bad {bad}/{err.numel()}"
          f" {'OK' if bad == 0 else 'FAIL'}", flush=True)
    return bad == 0


def diag_umma():
    ok = True
    for n, k in [(16, 64), (64, 64), (64, 128), (128, 64), (256, 64), (32, 32), (80, 96), (512, 256), (64, 576), (160, 1024)]:
        try:
            e = K.selftest_umma(n, k)
            good = e < 0.05
            ok &= good
            print(f"  selftest N={n:<4} K={k:<5} max_abs_err {e:.4e} {'OK' if good else 'FAIL'}", flush=True)
        except Exception as ex:  # noqa: BLE001
            ok = False
            print(f"  selftest N={n} K={k} raised: {ex}", flush=True)
    return ok


def conv_case(B, cin, cout, H, W, k, s, act=True, res=False, f32=False, in_pad=0, out_pad=0, tail_pad=0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(B, cin, H, W, generator=g).to(dev)
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    xb = torch.zeros(B, H, W, cin + in_pad, device=dev, dtype=torch.bfloat16)
    xb[..., in_pad:] = x.permute(0, 2, 3, 1)
    xin = xb.permute(0, 3, 1, 2)[:, in_pad:]
    wp, bp = K.pack_conv_weight(w, b)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    ob = torch.full((B, Ho, Wo, cout + out_pad + tail_pad), 7.0, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
    out = ob.permute(0, 3, 1, 2)[:, out_pad:out_pad + cout]
    r = None
    if res:
        r = torch.randn(B, cout, Ho, Wo, generator=g).to(dev).to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    K.conv2d(xin, wp, bp, cout, k, s, act, residual=r, out=out)
    torch.cuda.synchronize()
    ref = F.conv2d(xin.float(), w.to(torch.bfloat16).float(), b, stride=s, padding=k // 2)
    if act:
        ref = F.silu(ref)
    if res:
        ref = ref + r.float()
    name = f"conv B{B} {cin}->{cout} {H}x{W} k{k}s{s} act{int(act)} res{int(res)} f32{int(f32)} pad{in_pad}/{out_pad}"
    ok = report(name, out, ref, 2e-2)
    if out_pad:
        ok &= bool((ob[..., :out_pad] == 7.0).all())
    if tail_pad:
        ok &= bool((ob[..., out_pad + cout:] == 7.0).all())
    return ok


def diag_conv():
    ok = True
    cases = [
        dict(B=1, cin=64, cout=64, H=16, W=16, k=1, s=1),
        dict(B=2, cin=64, cout=64, H=16, W=16, k=3, s=1),
        dict(B=2, cin=64, cout=128, H=16, W=16, k=3, s=2),
        dict(B=2, cin=32, cout=32, H=40, W=40, k=3, s=1, res=True),
        dict(B=3, cin=128, cout=256, H=20, W=20, k=3, s=1),
        dict(B=2, cin=96, cout=64, H=32, W=32, k=1, s=1, in_pad=32, out_pad=64),
        dict(B=2, cin=64, cout=10, H=20, W=20, k=1, s=1, act=False, f32=True, out_pad=64, tail_pad=6),
        dict(B=2, cin=64, cout=64, H=20, W=20, k=1, s=1, act=False, f32=True),
        dict(B=1, cin=512, cout=512, H=20, W=20, k=3, s=1),
        dict(B=2, cin=256, cout=512, H=40, W=40, k=3, s=2),
        dict(B=1, cin=16, cout=16, H=32, W=32, k=3, s=1, res=True),
        dict(B=1, cin=8, cout=8, H=32, W=32, k=3, s=1),
        dict(B=1, cin=768, cout=256, H=40, W=40, k=1, s=1),
        dict(B=2, cin=80, cout=160, H=24, W=24, k=3, s=2),
        dict(B=1, cin=160, cout=320, H=13, W=17, k=3, s=1),
        dict(B=1, cin=48, cout=96, H=15, W=15, k=3, s=2),
        dict(B=4, cin=64, cout=64, H=160, W=160, k=3, s=1),
        dict(B=2, cin=1024, cout=512, H=20, W=20, k=1, s=1),
    ]
    for c in cases:
        try:
            ok &= conv_case(**c)
        except Exception as ex:  # noqa: BLE001
            ok = False
            print(f"  conv case {c} raised: {ex}", flush=True)
    return ok


def diag_decode():
    ok = True
    for imgsz, B in ((128, 2), (640, 2)):
        raw = recipe.synthetic_raw_maps(B, imgsz, 10, -10.0)
        ref = torch.from_numpy(decode_np.decode([r.numpy() for r in raw], [4.0, 8.0, 16.0, 32.0], 10))
        y = K.detect_decode([r.to(dev) for r in raw], [4, 8, 16, 32], 10)
        ok &= report(f"decode NCHW fp32 imgsz {imgsz} boxes", y[:, :4].cpu(), ref[:, :4], 1e-3)
        ok &= report(f"decode NCHW fp32 imgsz {imgsz} probs", y[:, 4:].cpu(), ref[:, 4:], 1e-5)
        for dt in (torch.float32, torch.bfloat16):
            lv, lv_ref = [], []
            for r in raw:
                Bq, no, H, W = r.shape
                buf = torch.zeros(Bq, H, W, 80, device=dev, dtype=dt)
                buf[..., :no] = r.to(dev).permute(0, 2, 3, 1).to(dt)
                lv.append(buf.permute(0, 3, 1, 2)[:, :no])
                lv_ref.append(buf[..., :no].permute(0, 3, 1, 2).float().cpu().numpy())
            ref2 = torch.from_numpy(decode_np.decode(lv_ref, [4.0, 8.0, 16.0, 32.0], 10))
            y2 = K.detect_decode(lv, [4, 8, 16, 32], 10)
            ok &= report(f"decode NHWC {dt} imgsz {imgsz} boxes", y2[:, :4].cpu(), ref2[:, :4], 1e-3)
            ok &= report(f"decode NHWC {dt} imgsz {imgsz} probs", y2[:, 4:].cpu(), ref2[:, 4:], 1e-5)
    return ok


NMS_CASES = {
    "default": dict(conf_thres=0.001, iou_thres=0.7, max_det=300),
    "multilabel": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True),
    "agnostic": dict(conf_thres=0.001, iou_thres=0.5, max_det=100, agnostic=True),
    "classes": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, classes=[1, 3, 7]),
    "maxnms": dict(conf_thres=0.001, iou_thres=0.7, max_det=300, max_nms=200),
    "predict": dict(conf_thres=0.25, iou_thres=0.45, max_det=300),
}


def nms_compare(name, y, kw, ref_out, ref_kept):
    out, counts, kept = K.nms(y.to(dev), kw["conf_thres"], kw["iou_thres"], max_det=kw["max_det"],
                              max_nms=kw.get("max_nms", 30000), agnostic=kw.get("agnostic", False),
                              multi_label=kw.get("multi_label", False), classes=kw.get("classes"))
    torch.cuda.synchronize()
    out, counts, kept = out.cpu().numpy(), counts.cpu().numpy(), kept.cpu().numpy()
    ok = True
    for b in range(y.shape[0]):
        n = int(counts[b])
        same_n = n == ref_out[b].shape[0]
        rows = same_n and np.array_equal(out[b, :n].view(np.uint32), ref_out[b].view(np.uint32))
        kp = same_n and np.array_equal(kept[b, :n], ref_kept[b])
        if not (same_n and rows and kp):
            ok = False
            msg = f"n {n} vs {ref_out[b].shape[0]} rows {rows} kept {kp}"
            if same_n and n:
                d = np.nonzero((out[b, :n] != ref_out[b]).any(1))[0]
                msg += f" first_bad_row {d[:3]} got {out[b, d[:1]]} ref {ref_out[b][d[:1]]}" if d.size else ""
                dk = np.nonzero(kept[b, :n] != ref_kept[b])[0]
                msg += f" first_bad_kept {dk[:3]} got {kept[b, dk[:3]]} ref {ref_kept[b][dk[:3]]}" if dk.size else ""
            print(f"  {name} image {b}: MISMATCH {msg}", flush=True)
    print(f"  {name:<58} {'OK' if ok else 'FAIL'} counts {counts[:4].tolist()}", flush=True)
    return ok


def diag_nms():
    ok = True
    for regime in ("sparse", "vallike", "dense"):
        g = np.load(ROOT / "tests" / "golden" / f"decode_nms_{regime}.npz")
        y = torch.from_numpy(g["y"])
        for case, kw in NMS_CASES.items():
            ref_out = [g[f"{case}_out{b}"] for b in range(y.shape[0])]
            ref_kept = [g[f"{case}_kept{b}"] for b in range(y.shape[0])]
            try:
                ok &= nms_compare(f"nms golden {regime}/{case}", y, kw, ref_out, ref_kept)
            except Exception as ex:  # noqa: BLE001
                ok = False
                print(f"  nms golden {regime}/{case} raised: {ex}", flush=True)
    # full-size: 34k anchors against the oracle restatement (itself pinned on the golden vectors)
    for mu, tag in ((-11.0, "sparse"), (-10.0, "vallike"), (-7.5, "dense")):
        raw = recipe.synthetic_raw_maps(2, 640, 10, mu)
        y = torch.from_numpy(decode_np.decode([r.numpy() for r in raw], [4.0, 8.0, 16.0, 32.0], 10))
        for case in ("default", "multilabel"):
            kw = NMS_CASES[case]
            t = time.time()
            ref_out, ref_kept = nms_np.non_max_suppression(y.numpy(), return_kept=True, **kw)
            dt = time.time() - t
            try:
                ok &= nms_compare(f"nms 34k {tag}/{case} (oracle {dt:.1f}s)", y, kw, ref_out, ref_kept)
            except Exception as ex:  # noqa: BLE001
                ok = False
                print(f"  nms 34k {tag}/{case} raised: {ex}", flush=True)
    return ok


def diag_aux():
    ok = True
    g = torch.Generator().manual_seed(0)
    # stem
    x = torch.rand(2, 3, 64, 96, generator=g).to(dev)
    w = (torch.randn(32, 3, 3, 3, generator=g) * 0.3).to(dev)
    b = torch.randn(32, generator=g).to(dev)
    out = K.stem_conv(x, w.reshape(32, 27).contiguous(), b)
    ok &= report("stem 3->32 64x96", out, F.silu(F.conv2d(x, w, b, stride=2, padding=1)), 2e-2)
    # sppf pool
    c = 64
    buf = torch.zeros(2, 20, 20, 4 * c, device=dev, dtype=torch.bfloat16)
    buf[..., :c] = torch.randn(2, 20, 20, c, generator=g).to(dev)
    v = buf.permute(0, 3, 1, 2)
    K.sppf_pool(v, c)
    y0 = v[:, :c].float()
    y1 = F.max_pool2d(y0, 5, 1, 2); y2 = F.max_pool2d(y1, 5, 1, 2); y3 = F.max_pool2d(y2, 5, 1, 2)
    ok &= report("sppf pool 20x20 c64", v[:, c:].float(), torch.cat((y1, y2, y3), 1), 0.0)
    # upsample
    xi = torch.randn(2, 10, 12, 128, generator=g).to(dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    ob = torch.zeros(2, 20, 24, 192, device=dev, dtype=torch.bfloat16)
    K.upsample2x(xi, out=ob.permute(0, 3, 1, 2)[:, :128])
    ok &= report("upsample2x into slice", ob.permute(0, 3, 1, 2)[:, :128].float(), F.interpolate(xi.float(), scale_factor=2.0), 0.0)
    # dwconv
    xi = torch.randn(2, 16, 16, 64, generator=g).to(dev).to(torch.bfloat16).permute(0, 3, 1, 2)
    w = (torch.randn(32, 2, 3, 3, generator=g) * 0.3).to(dev)
    b = torch.randn(32, generator=g).to(dev)
    o = K.dwconv3x3s2(xi, w, b)
    ok &= report("dwconv 64->32 s2", o, F.silu(F.conv2d(xi.float(), w, b, stride=2, padding=1, groups=32)), 2e-2)
    return ok


def diag_model():
    from drone_yolo_b200.nn.tasks import DetectionModel

    ok = True
    for tag in ("n_repvgg_128", "n_repvgg_sf_64", "n_p2_64", "s_repvgg_64"):
        g = np.load(ROOT / "tests" / "golden" / f"convstack_{tag}.npz")
        torch.manual_seed(int(g["model_seed"]))
        m = DetectionModel(str(g["yaml"]), nc=10, verbose=False)
        recipe.apply_recipe(m)
        m = m.eval().to(dev)
        x = recipe.images(int(g["B"]), int(g["imgsz"]), int(g["imgsz"])).to(dev)
        try:
            y, raw = m(x)
            torch.cuda.synchronize()
            for i, r in enumerate(raw):
                ok &= report(f"model {tag} raw{i}", r.cpu(), torch.from_numpy(g[f"raw{i}"].astype(np.float32)), 2e-2)
            ok &= report(f"model {tag} boxes (px)", y[:, :4].cpu(), torch.from_numpy(g["y"][:, :4]), 0.5)
            ok &= report(f"model {tag} probs", y[:, 4:].cpu(), torch.from_numpy(g["y"][:, 4:]), 2e-2)
        except Exception as ex:  # noqa: BLE001
            ok = False
            print(f"  model {tag} raised: {ex}", flush=True)
    return ok


FAMILIES = {"umma": diag_umma, "conv": diag_conv, "decode": diag_decode, "nms": diag_nms, "aux": diag_aux, "model": diag_model}

if __name__ == "__main__":
    which = sys.argv[1:] or ["all"]
    names = list(FAMILIES) if which == ["all"] else which
    print(torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0), flush=True)
    status = 0
    for n in names:
        print(f"== {n}", flush=True)
        t0 = time.time()
        good = FAMILIES[n]()
        print(f"== {n}: {'PASS' if good else 'FAIL'} ({time.time() - t0:.1f}s)", flush=True)
        status |= 0 if good else 1
    sys.exit(status)
