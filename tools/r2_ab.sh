#!/bin/bash
# Round-2 GPU visit: parity of the new conv mode / side lanes first, then same-box A/B of the plan options, then the whole GPU suite.
# usage: tools/r2_ab.sh <tag>
tag=${1:-r2a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "conv_vs_fp32 or engine_vs_cpu_oracle or model_vs_reference" > gpurun_out/${tag}_pytest_conv.log 2>&1
echo "conv/engine pytest exit $?"; tail -5 gpurun_out/${tag}_pytest_conv.log
one() {  # name, env...
  name=$1; shift
  out=$(env "$@" timeout 300 python bench.py --no-cpu-baseline --steps ${STEPS:-30} --warmup 5 2>gpurun_out/${tag}_bench_${name}.err | tail -1)
  echo "$out" > gpurun_out/${tag}_bench_${name}.json
  echo "$name: $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['ms_per_launch_group'],4), round(d['roofline']['frac'],4), round(d['e2e']['value']), d['clocks']['sm_mhz'])" 2>&1 | tail -1)"
}
for rep in 1 2; do
  one base${rep} DY_HEAD_LANES=0 DY_NO_PAIRED=1
  one paired${rep} DY_HEAD_LANES=0
  one lanes1_${rep} DY_HEAD_LANES=1
  one lanes2_${rep} DY_HEAD_LANES=2
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/${tag}_pytest.log
