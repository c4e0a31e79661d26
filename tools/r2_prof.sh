#!/bin/bash
# ncu launch list of one step (side lanes off so that launch order == plan order) -> gpurun_out/<tag>_launches.csv
tag=${1:-r2p}
export DY_HEAD_LANES=0
timeout 300 python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"; cat gpurun_out/${tag}_plain.log | tail -1
