#!/bin/bash
# Release-speed knock-out matrix: time a few layers with every compile-time knock-out build present under lib/
# (DY_CONV_KNOCKOUT_BUILD=<mask> python -m drone_yolo_b200.build; mask bits: 1 epilogue math+stores, 2 MMA, 4 A loads,
#  8 B loads, 16 SiLU without MUFU, 32 no TMA store).
mkdir -p gpurun_out
for lib in drone_yolo_b200/lib/libdroneyolo.so drone_yolo_b200/lib/libdroneyolo_k*.so; do
  [ -f $lib ] || continue
  echo "##### $lib"
  for pat in "$@"; do
    DY_LIB=$PWD/$lib timeout 120 python tools/bench_conv.py "$pat" 2>&1 | grep -v DY_CONV_DBG
  done
done
