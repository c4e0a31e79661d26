#!/bin/bash
# whole-step knock-outs: the conv-stack plan timed with each compile-time knock-out build of the conv kernel (results are wrong by
# construction; only the time matters).  Clocks sampled per run.
export DY_HEAD_LANES=0
for lib in drone_yolo_b200/lib/libdroneyolo.so drone_yolo_b200/lib/libdroneyolo_k*.so; do
  nvidia-smi --query-gpu=clocks.sm,power.draw --format=csv,noheader,nounits -lms 50 > /tmp/clk.txt &
  smi=$!
  r=$(DY_LIB=$PWD/$lib timeout 200 python tools/profile_step.py --micro-batch 64 --steps 100 | tail -1)
  kill $smi
  clk=$(sort -t, -k2 -n -r /tmp/clk.txt | head -20 | awk -F, '{c+=$1; p+=$2; n++} END {printf "sm %.0f MHz, %.0f W (mean of the 20 highest-power samples)", c/n, p/n}')
  echo "$(basename $lib): $r | $clk"
done
