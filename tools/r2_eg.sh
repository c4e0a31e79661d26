#!/bin/bash
# (round-2 experiment, measured neutral: build the stem variant with DY_EXTRA_NVCC_FLAGS=-DDY_STEM_EG=4 DY_LIB_TAG=stem4 python -m drone_yolo_b200.build)
# A/B of epilogue-group counts: parity of the touched kernels, then the step time (plan order, no side lanes) per variant, twice
tag=${1:-r2eg}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv or stem or engine_vs_cpu_oracle or model_vs_reference" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest(default) exit $?"; tail -2 gpurun_out/${tag}_pytest.log
DY_K32_EG4=1 DY_LIB=$PWD/drone_yolo_b200/lib/libdroneyolo_stem4.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv or stem or engine_vs_cpu_oracle or model_vs_reference" > gpurun_out/${tag}_pytest4.log 2>&1; echo "pytest(eg4+stem4) exit $?"; tail -2 gpurun_out/${tag}_pytest4.log
export DY_HEAD_LANES=0
for rep in 1 2; do
  echo "base:      $(DY_S2TAIL_EG3=0 timeout 200 python tools/profile_step.py --micro-batch 64 --steps 20 | tail -1)"
  echo "s2tail3:   $(timeout 200 python tools/profile_step.py --micro-batch 64 --steps 20 | tail -1)"
  echo "+k32eg4:   $(DY_K32_EG4=1 timeout 200 python tools/profile_step.py --micro-batch 64 --steps 20 | tail -1)"
  echo "+stem4:    $(DY_LIB=$PWD/drone_yolo_b200/lib/libdroneyolo_stem4.so timeout 200 python tools/profile_step.py --micro-batch 64 --steps 20 | tail -1)"
  echo "all:       $(DY_K32_EG4=1 DY_LIB=$PWD/drone_yolo_b200/lib/libdroneyolo_stem4.so timeout 200 python tools/profile_step.py --micro-batch 64 --steps 20 | tail -1)"
done
