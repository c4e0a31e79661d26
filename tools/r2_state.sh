#!/bin/bash
# Round-2 state check: full GPU parity suite, default bench line, ncu launch list of one step (plan order).
tag=${1:-r2s}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err; echo "bench exit $?"; tail -1 gpurun_out/${tag}_bench.log
export DY_HEAD_LANES=0
timeout 300 python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"; tail -1 gpurun_out/${tag}_plain.log
