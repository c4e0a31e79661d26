#!/bin/bash
# Release-speed knock-out matrix: time a few layers with every compile-time knock-out build
# (DY_CONV_KNOCKOUT_BUILD=<mask> python -m drone_yolo_b200.build; mask bits: 1 epilogue, 2 MMA, 4 A loads, 8 B loads).
mkdir -p gpurun_out
for k in "" _k1 _k2 _k4 _k12 _k15; do
  lib=drone_yolo_b200/lib/libdroneyolo$k.so
  [ -f $lib ] || continue
  echo "##### $lib"
  for pat in "$@"; do
    DY_LIB=$PWD/$lib timeout 120 python tools/bench_conv.py "$pat" 2>&1 | grep -v DY_CONV_DBG
  done
done
