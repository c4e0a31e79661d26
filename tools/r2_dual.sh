#!/bin/bash
# A/B: whole batch as one plan vs two half-batch plans on two streams (DY_DUAL_STREAM=1)
tag=${1:-r2dual}
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "validator" > gpurun_out/${tag}_pytest_val.log 2>&1; echo "validator pytest exit $?"; tail -3 gpurun_out/${tag}_pytest_val.log
DY_DUAL_STREAM=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "engine_vs_cpu_oracle or predict" > gpurun_out/${tag}_pytest.log 2>&1; echo "dual pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
for rep in 1 2; do
  for d in 0 1; do
    out=$(DY_DUAL_STREAM=$d timeout 300 python bench.py --no-cpu-baseline --no-eager --no-config4 --no-regimes --steps 30 --warmup 5 2>gpurun_out/${tag}_bench_$d.err | tail -1)
    echo "$out" > gpurun_out/${tag}_bench_${d}_${rep}.json
    echo "dual=$d: $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), 'plan', round(d['roofline']['ms_per_launch_group'],4), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['clocks']['sm_mhz'], d['gpu_launches'])" 2>&1 | tail -1)"
  done
done
