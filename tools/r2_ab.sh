#!/bin/bash
# Round-2 GPU visit: parity first, then same-box A/B of plan / kernel options (each a separate bench.py process).
# usage: tools/r2_ab.sh <tag> [--full] -- name1 "ENV=.. ENV=.." name2 "..."
tag=${1:-r2a}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv or engine_vs_cpu_oracle or model_vs_reference or tail" > gpurun_out/${tag}_pytest_conv.log 2>&1
echo "conv/engine pytest exit $?"; tail -4 gpurun_out/${tag}_pytest_conv.log
one() {  # name, env...
  name=$1; shift
  out=$(env "$@" timeout 300 python bench.py --no-cpu-baseline --steps ${STEPS:-30} --warmup 5 2>gpurun_out/${tag}_bench_${name}.err | tail -1)
  echo "$out" > gpurun_out/${tag}_bench_${name}.json
  echo "$name: $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch_group'],4), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
}
for rep in 1 2; do
  i=1
  while [ $i -le $# ]; do
    n=${!i}; i=$((i+1)); e=${!i}; i=$((i+1))
    one ${n}_${rep} $e
  done
done
