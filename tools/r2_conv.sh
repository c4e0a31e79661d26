#!/bin/bash
# parity of the conv kernel + single-layer timings of the layers the paired halo mode serves, with and without it
tag=${1:-r2c}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "conv_vs_fp32" > gpurun_out/${tag}_pytest_conv.log 2>&1
echo "conv pytest exit $?"; tail -3 gpurun_out/${tag}_pytest_conv.log
for f in "3x3 128->128" "3x3 256->128"; do
  echo "== generic"; DY_NO_PAIRED=1 timeout 120 python tools/bench_conv.py "$f" 2>&1 | tail -n +2
  echo "== paired"; timeout 120 python tools/bench_conv.py "$f" 2>&1 | tail -n +2
  for extra in "$@"; do
    [ "$extra" = "$tag" ] && continue
    echo "== paired $extra"; env $extra timeout 120 python tools/bench_conv.py "$f" 2>&1 | tail -n +2
  done
done
