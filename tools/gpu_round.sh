#!/bin/bash
# One GPU-box visit: parity tests, bench, ncu launch list and one full capture of the P2 head convs.
# usage: tools/gpu_round.sh <tag>   (outputs land in gpurun_out/<tag>_*)
tag=${1:-run}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.log 2>gpurun_out/${tag}_bench.err; echo "bench exit $?"; tail -1 gpurun_out/${tag}_bench.log
timeout 300 python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/${tag}_launches.csv python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_igemm -s 57 -c 3 \
    -f -o gpurun_out/${tag}_conv_full python tools/profile_step.py --micro-batch 64 > gpurun_out/${tag}_ncu2.log 2>&1
echo "full capture exit $?"
timeout 600 python tools/bench_decode_nms.py > gpurun_out/${tag}_config4.jsonl 2>gpurun_out/${tag}_config4.err; echo "config 4 exit $?"
timeout 300 python tools/bench_preprocess.py > gpurun_out/${tag}_preprocess.log 2>&1; echo "preprocess exit $?"; tail -2 gpurun_out/${tag}_preprocess.log
timeout 300 python tools/bench_slicer.py > gpurun_out/${tag}_slicer.jsonl 2>gpurun_out/${tag}_slicer.err &&
timeout 300 python tools/bench_slicer.py --conf 0.001 >> gpurun_out/${tag}_slicer.jsonl 2>>gpurun_out/${tag}_slicer.err; echo "slicer exit $?"; cat gpurun_out/${tag}_slicer.jsonl
