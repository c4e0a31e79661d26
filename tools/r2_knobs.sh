#!/bin/bash
# env-knob sweep on single layers (graph-replayed launches): usage tools/r2_knobs.sh "<layer pattern>" ... ; knob sets in KNOBS (';'-separated)
IFS=';' read -ra SETS <<< "${KNOBS:-;DY_CONV_NBUF2=1;DY_CONV_MAXBN=128;DY_CONV_MAXBN=128 DY_CONV_NBUF2=1}"
for pat in "$@"; do
  for ks in "${SETS[@]}"; do
    echo "== [$ks]"
    env $ks DY_CONV_VERBOSE=1 timeout 120 python tools/bench_conv.py "$pat" 2>&1 | grep -v DY_CONV_DBG | awk '/^conv k/ && !seen[$0]++ {print} !/^conv k/ {print}'
  done
done
