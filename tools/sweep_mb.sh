#!/bin/bash
mkdir -p gpurun_out
for mb in 4 8 16 32 64; do
  timeout 300 python bench.py --steps 10 --warmup 3 --micro-batch $mb --no-cpu-baseline > gpurun_out/bench_mb$mb.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_mb$mb.log").read().strip().splitlines()[-1])
    print("mb", $mb, "img/s", round(d["value"]), "ms", round(d["ms_per_step"],2), "conv_ms", round(d["stages"]["conv_stack_decode_ms"],2), "nms_ms", round(d["stages"]["nms_ms"],3), "TF", round(d["roofline"]["achieved"]), "e2e", round(d["e2e"]["value"]), d["clocks"])
except Exception as e:
    print("mb", $mb, "failed", e)
PY
done
