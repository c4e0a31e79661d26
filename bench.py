#!/usr/bin/env python
"""bench.py — images/sec of the Drone-YOLO inference hot path (conv stack + decode + NMS) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--scale s] [--imgsz 640] [--batch 64] [--micro-batch 0]

One "step" = one pass of the hot path over one synthetic batch (`--batch` images PER GPU: weak scaling).
Default workload = BASELINE.json configs[1]: Drone-YOLO-s, 640x640, batch 64, bf16 conv stack, 1xB200.
Prints ONE JSON line (see the task contract): `value` is device-timed throughput with inputs resident in HBM,
`e2e` includes the pinned-host -> device copy of every batch and the device -> host read of the detections.
`--impl reference` times the CPU port of the reference path (oracle/, torch fp32 + numpy NMS) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

GFLOP_PER_IMAGE_640 = {"n": 12.193, "s": 36.655, "l": 204.80, "x": 316.160, "m": 97.97}   # SURVEY.md §8(d), deploy form
METRIC = "images/sec at 640px (conv+decode+NMS)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "cpu-worker"])
    ap.add_argument("--threads", type=int, default=0, help="cpu-worker: torch threads (0 = the reference's policy min(8, cores - 1))")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per CPU-baseline step")
    ap.add_argument("--no-config4", action="store_true", help="skip the isolated decode + NMS stage block (BASELINE config 4)")
    ap.add_argument("--no-eager", action="store_true", help="skip the same-box torch-eager GPU baseline")
    ap.add_argument("--no-regimes", action="store_true", help="skip the extra class-bias regimes of the NMS stage")
    ap.add_argument("--scale", default="s")
    ap.add_argument("--imgsz", type=int, default=640)
    ap.add_argument("--batch", type=int, default=64, help="images per GPU per step")
    ap.add_argument("--micro-batch", type=int, default=0)
    ap.add_argument("--conf", type=float, default=0.001)
    ap.add_argument("--iou", type=float, default=0.7)
    ap.add_argument("--max-det", type=int, default=300)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cls-delta", type=float, default=2.5,
                    help="class-bias shift of the seeded weight recipe (SURVEY.md 8(d)): the random-init class logits hardly vary, so "
                         "whole levels pass conf 0.001 together: 2.5 -> the 2000 P4+P5 anchors (5.9 %%), 2.65 -> 8400 (24.7 %%), "
                         "4.0 -> all 34 000 (dense, max_nms truncation)")
    return ap.parse_args()


def workload_name(a, n):
    return (f"Drone-YOLO-{a.scale} {a.imgsz}x{a.imgsz} batch {a.batch} per GPU bf16 end-to-end inference "
            f"(conv+decode+NMS) on {n}xB200, nc=10, conf {a.conf}, iou {a.iou}, max_det {a.max_det}")


def build_model(scale, cls_delta=2.5):
    import torch
    from drone_yolo_b200.nn.tasks import DetectionModel
    from oracle import recipe      # seeded test weights only (bench may use oracle/ for inputs + the CPU baseline)

    torch.manual_seed(0)
    m = DetectionModel(f"yolov8{scale}-p2-repvgg.yaml", nc=10, verbose=False)
    recipe.apply_recipe(m, cls_delta=cls_delta)
    return m.eval()


def algorithmic_gflop(scale, imgsz):
    return GFLOP_PER_IMAGE_640[scale] * (imgsz / 640.0) ** 2


# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "10",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------------------
def reference_threads():
    """The reference's thread policy: NUM_THREADS = min(8, max(1, os.cpu_count() - 1)) (ultralytics/utils/__init__.py:44), applied
    with torch.set_num_threads (utils/torch_utils.py:229)."""
    return min(8, max(1, (os.cpu_count() or 1) - 1))


def cpu_port_run(a, batch, steps, warmup, threads, budget_s=150.0):
    """Time the CPU port of the reference path (oracle/torch_ref: the aten ops the reference's modules execute, fp32, Conv+BN
    folded and RepVGG blocks two-branch as BaseModel.fuse leaves them; oracle/decode_np + nms_np) on `batch` synthetic images per
    step.  Returns per-step seconds split into conv stack / decode / NMS."""
    import torch
    from oracle import decode_np, nms_np, recipe, torch_ref

    torch.set_num_threads(threads)
    m = build_model(a.scale, a.cls_delta)
    torch_ref.fuse_like_reference(m)
    det = m.model[-1]
    strides = [float(s) for s in det.stride.tolist()]
    x = recipe.images(batch, a.imgsz, a.imgsz)
    rows = []
    t_begin = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        raw = torch_ref.forward_raw(m, x)
        t1 = time.perf_counter()
        y = decode_np.decode([r.numpy() for r in raw], strides, det.nc)
        t2 = time.perf_counter()
        out = nms_np.non_max_suppression(y, conf_thres=a.conf, iou_thres=a.iou, max_det=a.max_det)
        t3 = time.perf_counter()
        if i >= warmup:
            rows.append((t1 - t0, t2 - t1, t3 - t2))
        if i >= warmup and time.perf_counter() - t_begin > budget_s:
            break
    n = len(rows)
    conv, dec, nms = (sum(r[k] for r in rows) / n for k in range(3))
    cand = float((y[:, 4:].max(1) > a.conf).sum()) / batch
    return {"images_per_s": batch / (conv + dec + nms), "steps": n, "batch": batch, "threads": torch.get_num_threads(),
            "conv_s": conv, "decode_s": dec, "nms_s": nms, "candidates_per_image": cand,
            "kept_per_image": sum(o.shape[0] for o in out) / batch}


def run_cpu_worker(a):
    """One CPU-baseline measurement in its own process (fresh thread pools): prints one JSON object."""
    r = cpu_port_run(a, a.cpu_batch, max(1, a.steps), max(0, a.warmup), a.threads or reference_threads())
    r["cls_delta"] = a.cls_delta
    print(json.dumps(r), flush=True)


def cpu_worker_subprocess(a, threads, cls_delta, steps=5, warmup=2, timeout=240):
    cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "cpu-worker", "--threads", str(threads), "--cls-delta", str(cls_delta),
           "--cpu-batch", str(a.cpu_batch), "--steps", str(steps), "--warmup", str(warmup), "--scale", a.scale, "--imgsz", str(a.imgsz),
           "--conf", str(a.conf), "--iou", str(a.iou), "--max-det", str(a.max_det)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env)
    if p.returncode != 0:
        raise RuntimeError(p.stderr[-300:])
    return json.loads(p.stdout.strip().splitlines()[-1])


def cpu_sample_text(r):
    return (f"{r['steps']} steps x {r['batch']} images of the same workload after 2 warm-up steps, {r['threads']} torch threads, in a "
            f"subprocess: conv stack {r['conv_s']:.3f} s + decode {r['decode_s']:.3f} s + NMS {r['nms_s']:.3f} s per step "
            f"({r['candidates_per_image']:.0f} candidates per image); CPU port of the reference path (torch fp32 aten ops + numpy "
            f"decode / NMS: the reference tree does not travel to the GPU box)")


def run_reference(a):
    """The reference arm: the CPU port of the reference's predict path on the host cores with the reference's own thread
    policy, B = --cpu-batch images per step, K steps after W warm-up steps (bounded to a few minutes); an all-cores row rides
    along.  kind = "port": /root/reference (Python) cannot travel to the GPU box."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = reference_threads()
    r = cpu_worker_subprocess(a, threads, a.cls_delta, steps=max(1, a.steps), warmup=max(0, min(a.warmup, 2)), timeout=400)
    extra = {}
    try:
        cores = os.cpu_count() or 1
        if cores != threads:
            ra = cpu_worker_subprocess(a, cores, a.cls_delta, steps=5, warmup=2)
            extra["all_cores"] = {"value": ra["images_per_s"], "unit": "images/s", "cores": ra["threads"], "sample": cpu_sample_text(ra)}
    except Exception as ex:  # noqa: BLE001
        extra["all_cores"] = {"value": None, "error": str(ex)[:200]}
    value = r["images_per_s"]
    sample = cpu_sample_text(r)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": a.gpus, "steps": r["steps"],
        "warmup": max(0, min(a.warmup, 2)), "ms_per_step": 1e3 * r["batch"] / value, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a, a.gpus), "sample": sample, "kind": "port", "thread_policy": "min(8, cores - 1) (ultralytics/utils/__init__.py:44)",
                   "weights": f"random-init (seed 0) + seeded BN recipe, class-bias shift {a.cls_delta}"},
        "cpu_baseline": dict({"value": value, "unit": "images/s", "cores": r["threads"], "kind": "port", "sample": sample,
                              "split_s_per_step": {"conv": r["conv_s"], "decode": r["decode_s"], "nms": r["nms_s"]}}, **extra),
        "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
def run_ours(a):
    import copy
    import importlib.util

    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from drone_yolo_b200 import _C, YOLO
    from drone_yolo_b200.engine.engine import Engine
    from drone_yolo_b200.parallel import DetectionGather
    from oracle import recipe

    B = a.batch
    model_cpu = build_model(a.scale, a.cls_delta)
    model = copy.deepcopy(model_cpu).to(dev).fuse(verbose=False)
    eng = Engine(model, B, a.imgsz, dev, micro_batch=a.micro_batch, conf=a.conf, iou=a.iou, max_det=a.max_det,
                 cuda_graph=not a.no_graph)
    host = recipe.images(B, a.imgsz, a.imgsz, seed=2 + rank).pin_memory()
    eng.images.copy_(host)
    gather = DetectionGather(B, a.max_det, dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(v):
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        out, counts = eng.step()
        if world > 1:
            gather.gather(out, counts)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def graph_of(fn):
        fn(); torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(3):
            g.replay()
        return g

    # ---- `value`: the whole step on a batch resident in HBM (device-timed, max over ranks) ----------------------------
    sampler = ClockSampler(local)           # nvidia-smi answers every ~100 ms: sample from the warm-up on (same load) to the end
    if rank == 0:                           # of the timed region so that a 20-step run still yields several readings
        sampler.start()
    for _ in range(max(a.warmup, 3)):
        step_resident()
    ms_total = timed(step_resident, a.steps)
    t_fill = time.perf_counter()
    while time.perf_counter() - t_fill < 0.3:   # K steps may last less than a sampler period: keep the very same load
        step_resident()                         # running (untimed) for a few more periods
    torch.cuda.synchronize(dev)
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "warm-up + timed region + 0.3 s of the same load after it"

    # ---- `e2e`: the call a user makes.  YOLO.predict(stream of host frames, stream=True): every batch of B x world uint8 HWC
    # BGR frames (pinned host memory) is sharded over the ranks, uploaded (H2D inside the timed region), letterboxed on the
    # GPU, run through the engine, gathered to rank 0, read back (D2H) and turned into Results there.  Wall clock.
    host8 = (host * 255.0).round().to(torch.uint8)
    frames_t = host8.permute(0, 2, 3, 1).flip(-1).contiguous().pin_memory()        # (B, H, W, 3) BGR, pinned
    frames = [f.numpy() for f in frames_t]                                           # views of the pinned block
    frames_global = frames * world                                                   # a rank only ever touches its own shard
    yolo = YOLO(model)
    kw = dict(imgsz=a.imgsz, conf=a.conf, iou=a.iou, max_det=a.max_det, device=dev, micro_batch=a.micro_batch,
              cuda_graph=not a.no_graph, batch=B * world, stream=True)

    def frame_stream(nb):
        for _ in range(nb):
            yield from frames_global

    def predict_run(nb):
        n = det = 0
        for r in yolo.predict(frame_stream(nb), **kw):
            n += 1
            det += len(r)
        return n, det

    predict_run(max(a.warmup, 3))                                                    # builds the engine, captures its graphs
    barrier()
    t0 = time.perf_counter()
    n_res, n_det = predict_run(a.steps)
    torch.cuda.synchronize(dev)
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    barrier()
    assert rank != 0 or n_res == a.steps * B * world, (n_res, a.steps * B * world)

    # ---- engine-level pipeline (no Results, no letterbox): pinned uint8 NCHW -> slot, step, D2H; explains the predictor's overhead
    eng8 = yolo.predictor.engine_for(B, a.imgsz, a.imgsz, torch.uint8)
    host8p = host8.pin_memory()
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    slot_free = [torch.cuda.Event() for _ in range(2)]
    d2h_done = [torch.cuda.Event() for _ in range(2)]
    out_hosts = [torch.empty((world * B, a.max_det, 6), dtype=torch.float32).pin_memory() for _ in range(2)]
    cnt_hosts = [torch.empty((world * B,), dtype=torch.int32).pin_memory() for _ in range(2)]
    st = {"i": 0, "primed": False}

    def enqueue_h2d(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(slot_free[i % 2])
            eng8.image_slots[i % 2].copy_(host8p, non_blocking=True)
            h2d_done[i % 2].record(copy_stream)

    def step_engine_e2e():
        i = st["i"]
        cur = torch.cuda.current_stream(dev)
        if not st["primed"]:
            for ev in slot_free:
                ev.record(cur)
            enqueue_h2d(i)
            st["primed"] = True
        enqueue_h2d(i + 1)
        cur.wait_event(h2d_done[i % 2])
        out, counts = eng8.step(slot=i % 2)
        slot_free[i % 2].record(cur)
        oa, ca = gather.gather(out, counts)
        if rank == 0:
            out_hosts[i % 2].copy_(oa, non_blocking=True)
            cnt_hosts[i % 2].copy_(ca, non_blocking=True)
        d2h_done[i % 2].record(cur)
        if i > 0:
            d2h_done[(i - 1) % 2].synchronize()
        st["i"] = i + 1

    with torch.inference_mode():                 # the predictor built this engine (and its slots) under inference_mode
        for _ in range(2):
            step_engine_e2e()
        ms_eng_e2e = timed(step_engine_e2e, a.steps)

    # ---- per-stage device times for the roofline: conv-stack plan alone, NMS alone (same buffers, separate graphs) ----
    g_plan = graph_of(lambda: eng.enqueue(nms=False))
    g_nms = graph_of(lambda: _C.check(_C.lib().dy_program_run(eng._nms_prog, 0, 0, _C.stream_ptr(dev)), "nms"))
    ms_plan = timed(g_plan.replay, a.steps) / a.steps
    ms_nms = timed(g_nms.replay, a.steps) / a.steps
    cand0 = float((eng.y[:, 4:].amax(1) > a.conf).sum()) / B

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    tf_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json, sustained)" if peaks else "fallback (B200_PROFILING.md)"

    # DRAM traffic of the launch group: measured separately under ncu (a number taken under a profiler is never a bench
    # value, but the byte counters are exact) and committed under profiles/; only valid for the default workload
    traffic, traffic_src = None, None
    for name in ("r02_dram_traffic_s640_b64.json", "r01_dram_traffic_s640_b64.json"):
        tf = ROOT / "profiles" / name
        if tf.exists() and (a.scale, a.imgsz, B) == ("s", 640, 64):
            t = json.loads(tf.read_text())
            traffic, traffic_src = t["plan_dram_bytes"], t["source"]
            break
    imgs = B * world
    ms_step = ms_total / a.steps
    value = imgs / (ms_step / 1e3)
    e2e_value = imgs * a.steps / dt_e2e
    gf = algorithmic_gflop(a.scale, a.imgsz)
    conv_tflops = gf * B / ms_plan                       # GFLOP / ms == TFLOP/s
    nms_bytes = B * (eng.A * (4 + eng.nc) * 4 + a.max_det * 24 + 4)
    dec_bytes = B * eng.A * (eng.model.model[-1].no * 4 + (4 + eng.nc) * 4)   # fp32 raw maps in this build
    d2h = (world * B * a.max_det * 6 + world * B) * 4
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": workload_name(a, world), "global_batch": imgs, "micro_batch": eng.mb, "parallelism": f"dp{world}",
                   "cuda_graph": not a.no_graph,
                   "decode": ("fused into the Detect conv tails for %d of %d levels (logits never reach HBM); standalone kernel for the rest"
                              % (sum(r is None for r in eng.plan.raw_refs), len(eng.plan.raw_refs))),
                   "weights": f"random-init (seed 0) + seeded BN recipe, class-bias shift {a.cls_delta} ({cand0:.0f} of {eng.A} anchors per image pass conf)",
                   "l2": f"inputs larger than L2: {B * 3 * a.imgsz * a.imgsz * 4 / 1e6:.0f} MB of images per step, "
                         f"arena {eng.plan.arena_bytes / 1e6:.0f} MB per micro-batch"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": frames_t.numel() * world + imgs * 32,
                "d2h_bytes_per_step": d2h, "ms_per_step": dt_e2e / a.steps * 1e3, "clock": "host wall clock, max over ranks",
                "call": f"YOLO.predict(stream of uint8 {a.imgsz}x{a.imgsz}x3 BGR host frames, stream=True, batch={imgs}) -> Results: "
                        f"{a.steps} batches after {max(a.warmup, 3)} warm-up batches; every batch is uploaded from pinned host memory, letterboxed on "
                        "the GPU, run (conv+decode+NMS+rescale), read back and turned into Results on rank 0",
                "detections_per_image": n_det / max(n_res, 1),
                "engine_level": {"value": imgs / (ms_eng_e2e / a.steps / 1e3), "unit": "images/s",
                                 "what": "the same two-slot pipeline driven through Engine.step directly (pinned uint8 NCHW batch, no letterbox, no Results)"}},
        "gpu_launches": eng.launches_per_step * a.steps,
        "clocks": clocks,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (conv stack plan: stem + tcgen05 convs + pool + fused decode)",
                     "achieved": conv_tflops, "peak": tf_peak, "unit": "TFLOP/s", "frac": conv_tflops / tf_peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "ms_per_launch_group": ms_plan,
                     "algorithmic_gflop_per_image": gf},
        "stages": {"conv_stack_decode_ms": ms_plan, "nms_ms": ms_nms, "nms_candidates_per_image": cand0,
                   "nms_hbm_gbs": nms_bytes / ms_nms / 1e6, "nms_frac_of_hbm": nms_bytes / ms_nms / 1e6 / hbm_peak,
                   "decode_algorithmic_bytes": dec_bytes},
    }

    def guarded(key, fn, where=None):
        try:
            (where if where is not None else line)[key] = fn()
        except Exception as ex:  # noqa: BLE001
            (where if where is not None else line)[key] = {"error": f"{type(ex).__name__}: {str(ex)[:200]}"}

    if world == 1 and not a.no_regimes:
        # the same step with more anchors passing conf (the recipe's class-bias shift): NMS work depends on it, the conv stack does not
        def regimes():
            rows = []
            for delta in (2.65, 4.0):
                m2 = build_model(a.scale, delta).to(dev).fuse(verbose=False)
                e2 = Engine(m2, B, a.imgsz, dev, micro_batch=a.micro_batch, conf=a.conf, iou=a.iou, max_det=a.max_det, cuda_graph=not a.no_graph)
                e2.images.copy_(host)
                for _ in range(3):
                    e2.step()
                ms2 = timed(lambda: e2.step(), a.steps) / a.steps
                g2 = graph_of(lambda: _C.check(_C.lib().dy_program_run(e2._nms_prog, 0, 0, _C.stream_ptr(dev)), "nms"))
                msn = timed(g2.replay, a.steps) / a.steps
                rows.append({"cls_delta": delta, "candidates_per_image": float((e2.y[:, 4:].amax(1) > a.conf).sum()) / B,
                             "ms_per_step": ms2, "images_per_s": B / ms2 * 1e3, "nms_ms": msn,
                             "nms_frac_of_hbm": nms_bytes / msn / 1e6 / hbm_peak})
                del e2, g2, m2
                torch.cuda.empty_cache()
            return rows
        guarded("regimes", regimes, line["stages"])

    if world == 1 and not a.no_config4:
        # BASELINE config 4 inside the driver-run line: isolated decode + NMS, B = 256, 34 000 / 136 000 anchors, its own clock sample
        def config4():
            spec = importlib.util.spec_from_file_location("bench_decode_nms", ROOT / "tools" / "bench_decode_nms.py")
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            cs = ClockSampler(local)
            cs.start()
            rows = mod.run_config4(dev, 256, (640, 1280), iters=10, layouts=False, multi_label=True)
            ck = cs.stop()
            return {"what": "isolated Detect decode + NMS, B=256, nc=10, conf 0.001, iou 0.7, max_det 300; ALGORITHMIC bytes over CUDA-event time "
                            "(10 launches after 3 warm-ups, inputs far larger than L2) against MEASURED_PEAKS.json's copy bandwidth",
                    "rows": rows, "clocks": ck}
        guarded("config4", config4, line["stages"])

    if world == 1 and not a.no_eager:
        # the same-box GPU bar: torch eager + cuDNN + torchvision.ops.nms on the same weights and images
        def eager():
            from oracle import eager_gpu
            rows = []
            cs = ClockSampler(local)
            cs.start()
            for dtype, deploy in ((torch.bfloat16, False), (torch.float16, False), (torch.bfloat16, True)):
                m2 = eager_gpu.prepare(model_cpu, dev, dtype, deploy)
                x32 = eng.images
                for _ in range(3):
                    eager_gpu.step(m2, x32, a.conf, a.iou, a.max_det, dtype)
                torch.cuda.synchronize(dev)
                k = max(3, min(a.steps, 10))
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                conv_ms = nms_ms = 0.0
                for _ in range(k):
                    e[0].record()
                    yv = eager_gpu.forward(m2, x32.to(dtype).contiguous(memory_format=torch.channels_last))
                    e[1].record()
                    outv = eager_gpu.non_max_suppression(yv, a.conf, a.iou, max_det=a.max_det)
                    e[2].record()
                    torch.cuda.synchronize(dev)
                    conv_ms += e[0].elapsed_time(e[1]) / k
                    nms_ms += e[1].elapsed_time(e[2]) / k
                rows.append({"dtype": str(dtype).replace("torch.", ""), "fuse": "deploy (RepVGG merged)" if deploy else "reference (Conv+BN folded, RepVGG two-branch)",
                             "images_per_s": B / (conv_ms + nms_ms) * 1e3, "conv_decode_ms": conv_ms, "nms_ms": nms_ms, "steps": k,
                             "kept_per_image": sum(o.shape[0] for o in outv) / B})
                del m2
                torch.cuda.empty_cache()
            ck = cs.stop()
            best = max(r["images_per_s"] for r in rows)
            return {"value": best, "unit": "images/s", "kind": "port",
                    "what": "torch eager on the same GPU, weights and resident images: channels_last cuDNN conv stack (module walk of oracle/torch_ref), "
                            "aten decode, ops.non_max_suppression restated with torchvision.ops.nms per image (oracle/eager_gpu.py); best row",
                    "cudnn": torch.backends.cudnn.version(), "rows": rows, "clocks": ck}
        guarded("gpu_eager_baseline", eager)

    if world == 1 and not a.no_cpu_baseline:
        def cpu():
            t = reference_threads()
            r = cpu_worker_subprocess(a, t, a.cls_delta)
            d = {"value": r["images_per_s"], "unit": "images/s", "cores": r["threads"], "kind": "port", "sample": cpu_sample_text(r),
                 "split_s_per_step": {"conv": r["conv_s"], "decode": r["decode_s"], "nms": r["nms_s"]},
                 "thread_policy": "the reference's: min(8, cores - 1) (ultralytics/utils/__init__.py:44)"}
            try:
                r2 = cpu_worker_subprocess(a, t, 2.65)
                d["regime_25pct"] = {"value": r2["images_per_s"], "cls_delta": 2.65, "candidates_per_image": r2["candidates_per_image"],
                                     "split_s_per_step": {"conv": r2["conv_s"], "decode": r2["decode_s"], "nms": r2["nms_s"]}}
                cores = os.cpu_count() or 1
                if cores != t:
                    r3 = cpu_worker_subprocess(a, cores, a.cls_delta)
                    d["all_cores"] = {"value": r3["images_per_s"], "cores": r3["threads"],
                                      "split_s_per_step": {"conv": r3["conv_s"], "decode": r3["decode_s"], "nms": r3["nms_s"]}}
            except Exception as ex:  # noqa: BLE001
                d["extra_rows_error"] = str(ex)[:200]
            return d
        guarded("cpu_baseline", cpu)
        if "error" in line["cpu_baseline"]:
            line["cpu_baseline"] = {"value": None, "unit": "images/s", "cores": 0, "kind": "port", "sample": "failed: " + line["cpu_baseline"]["error"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "cpu-worker":
        run_cpu_worker(args)
    else:
        run_ours(args)
