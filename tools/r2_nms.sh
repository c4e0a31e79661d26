#!/bin/bash
tag=${1:-r2n}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "nms or predict or engine_vs_cpu or slicer" > gpurun_out/${tag}_pytest_nms.log 2>&1
echo "nms pytest exit $?"; tail -5 gpurun_out/${tag}_pytest_nms.log
for L in prev cur prev cur; do
  if [ $L = prev ]; then export DY_LIB=drone_yolo_b200/lib/libdroneyolo_prev.so; else unset DY_LIB; fi
  echo "== $L"; timeout 300 python tools/bench_decode_nms.py --imgsz 640 1280 2>&1 | grep '"nms"' | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['imgsz'], d['regime'], 'ml' if d['multi_label'] else 'sl', d['candidates_per_image'], d['ms'], d['frac_of_measured_hbm'])"
done
