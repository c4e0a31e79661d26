"""Per-launch roofline table: joins the symbolic layer plan (CPU, no GPU needed) with an ncu launch list
(`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv python tools/profile_step.py ...`).

    python tools/layer_table.py gpurun_out/launches.csv [--scale s] [--imgsz 640] [--batch 64] [--md out.md]

For every launch: GEMM shape, algorithmic FLOPs and ideal bytes (read input slice once + weights + write output
[+ residual]), time, achieved TFLOP/s and GB/s, and the roofline time max(flop/peak_tc, bytes/peak_hbm) using
MEASURED_PEAKS.json (sustained bf16, copy bandwidth).  Launch order == op order of the plan.
"""
import argparse
import csv
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def plan_ops(scale, imgsz, batch):
    import torch.nn as nn
    from bench import build_model
    from drone_yolo_b200.engine.plan import LayerPlan
    model = build_model(scale).fuse(verbose=False)
    lp = LayerPlan.__new__(LayerPlan)
    lp.model, lp.mb, lp.H, lp.W, lp.device = model, batch, imgsz, imgsz, None
    lp.bufs, lp.ops, lp.keep = [], [], []
    lp.fuse_upsample = True
    lp.fuse_tail = True
    lp.fuse_decode = True
    lp.fuse_cv1 = True
    lp.split_up, lp.up_split = True, {}
    lp.head_lanes, lp.lane = 0, 0
    lp._build_symbolic()
    rows = []
    for op in lp.ops:
        kind = op["kind"]
        if kind == "conv":
            if "mod" in op:
                cout, k, s = lp._conv_geom(op["mod"])
            else:
                cout, k, s = op["cout"], op["k"], op["s"]
            inp, out = op["inp"], op["out"]
            cin = inp.c
            tail = op.get("tail")
            dec = bool(tail) and len(tail) > 4
            if out is None:
                out = tail[3] if not dec else inp
            M = batch * out.H * out.W
            flops = 2.0 * M * cout * k * k * cin
            oes = 4 if dec else lp.bufs[out.buf].esz
            # fused decode writes 4 box values (box tail) or nc class scores (class tail) per anchor instead of the raw map
            ocols = (4 if tail[4][0] == 1 else lp.model.model[-1].nc) if dec else (tail[2] if tail else cout)
            byts = batch * inp.H * inp.W * cin * 2 + M * ocols * oes + k * k * cin * cout * 2
            if tail:
                flops += 2.0 * M * cout * tail[2]
            if op.get("res") is not None:
                byts += M * cout * 2
            if op.get("up") is not None:
                byts += 4 * M * cout * 2
            if op.get("pre") is not None:
                byts += (M // 4) * cout * 4                   # half-resolution fp32 addend, read once
            rows.append(dict(name=f"conv{k}x{k}s{s} {cin}->{cout} @{out.H}" + (" +res" if op.get("res") is not None else "") + (" +up2x" if op.get("up") is not None else "") + (" +pre(up)" if op.get("pre") is not None else "") + (f" +1x1->{op['tail'][2]}" if op.get("tail") else "") + (" +decode" if dec else "") + (" f32" if oes == 4 else ""),
                             flops=flops, bytes=byts))
        elif kind == "stem":
            out = op["out"]
            M = batch * out.H * out.W
            rows.append(dict(name=f"stem 3->{out.c} @{out.H}", flops=2.0 * M * out.c * 27, bytes=batch * 3 * imgsz * imgsz * 1 + M * out.c * 2))
        elif kind == "pool":
            inp = op["inp"]
            rows.append(dict(name=f"sppf pool c={op['c']} @{inp.H}", flops=0, bytes=batch * inp.H * inp.W * op["c"] * 2 * 4))
        elif kind == "upsample":
            inp = op["inp"]
            rows.append(dict(name=f"upsample c={inp.c} @{inp.H}", flops=0, bytes=batch * inp.H * inp.W * inp.c * 2 * 5))
        elif kind == "dwconv":
            inp, out = op["inp"], op["out"]
            rows.append(dict(name=f"dwconv {inp.c}->{out.c} @{out.H}", flops=2.0 * batch * out.H * out.W * out.c * 18,
                             bytes=batch * (inp.H * inp.W * inp.c + out.H * out.W * out.c) * 2))
        elif kind == "decode":
            A = sum(r.H * r.W for r in op["levels"])
            nc = op["det"].nc
            rows.append(dict(name=f"decode A={A}", flops=0, bytes=batch * A * ((64 + nc + (-nc) % 16) * 4 + (4 + nc) * 4)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--scale", default="s")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--md", default="")
    a = ap.parse_args()
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    tc = peaks.get("bf16_tflops_sustained", 1397.2) * 1e12
    bw = peaks.get("hbm_gbs", 6553.6) * 1e9
    first = open(a.csv).read(4096)
    if '"Kernel Name"' in first or first.lstrip().startswith(('"', "==")):
        rows = list(csv.reader(l for l in open(a.csv) if l.startswith('"')))
        hdr = rows[0]
        ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
        launches = [(r[ki].split("(")[0].replace("void ", ""), float(r[vi].replace(",", "")) / 1e3) for r in rows[1:] if r[mi] == "gpu__time_duration.sum"]
    else:
        # the text of Engine.profile(verbose=True) / tools/profile_layers.py: warm per-op device times, "<us> us  <op>"
        import re
        launches = []
        for line in open(a.csv):
            m = re.match(r"\s*([\d.]+) us  (.*)", line)
            if m and "total (" not in m.group(2):
                launches.append(("(Engine.profile)" if not m.group(2).startswith("nms") else "nms_filter + nms_select", float(m.group(1))))
    ops = plan_ops(a.scale, a.imgsz, a.batch)
    # the NMS kernels follow the plan
    out = []
    out.append(f"| # | kernel | layer | us | TFLOP/s | GB/s (ideal bytes) | roofline us | x roofline |")
    out.append("|---|---|---|---|---|---|---|---|")
    tot_t = tot_r = 0.0
    for i, (kname, us) in enumerate(launches):
        if i < len(ops):
            o = ops[i]
            roof = max(o["flops"] / tc, o["bytes"] / bw) * 1e6
            out.append(f"| {i} | {kname} | {o['name']} | {us:.1f} | {o['flops'] / us / 1e6:.0f} | {o['bytes'] / us / 1e3:.0f} | {roof:.1f} | {us / roof:.2f} |")
            tot_t += us; tot_r += roof
        else:
            out.append(f"| {i} | {kname} | (post) | {us:.1f} | | | | |")
    out.append(f"\nplan total {tot_t:.0f} us, sum of per-layer rooflines {tot_r:.0f} us ({tot_t / tot_r:.2f}x)")
    text = "\n".join(out)
    print(text)
    if a.md:
        Path(a.md).write_text(text + "\n")


if __name__ == "__main__":
    main()
