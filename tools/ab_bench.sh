#!/bin/bash
# A/B on one box: lib/libdroneyolo_prev.so (built from another commit) vs the current library, alternating runs.
for i in 1 2 3; do for L in prev cur; do
  if [ $L = prev ]; then export DY_LIB=drone_yolo_b200/lib/libdroneyolo_prev.so; else unset DY_LIB; fi
  echo "$L: $(timeout 300 python bench.py --steps ${STEPS:-30} --warmup 5 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), d['clocks']['sm_mhz'])")"
done; done
