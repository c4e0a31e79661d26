#!/bin/bash
# compute-sanitizer pass over the warp-specialised kernels: memcheck, racecheck, synccheck and initcheck on a
# reduced set of the GPU parity tests (every conv mode incl. halo / paired / stride-2, fused Detect tails, SiLU tail, fused
# upsample, stem, pool, decode, NMS, tile merge).  Logs -> gpurun_out/<tag>_sanitize_<tool>.log; summary on stdout.
# usage: tools/sanitize.sh <tag> [tools...]
tag=${1:-san}; shift
tools=${@:-memcheck racecheck synccheck}
mkdir -p gpurun_out
SEL='umma_selftest or (conv_vs_fp32_reference and (2x64x64x16x16x3x1x or 2x64x128x16x16x3x2 or 2x32x32x40x40 or 3x128x256x20x20 or 2x96x64x32x32 or 2x64x10x20x20 or 1x160x320x13x17 or 2x64x64x48x48 or 4x128x128x40x40 or 16x128x64x48x32 or 2x16x32x64x96 or 2x1024x512x20x20 or 2x128x128x24x24)) or conv_fused_upsample or conv_fused_detect_tail or conv_fused_silu_tail or (stem_tensor_core and 2-320-256) or stem_pool_upsample or decode_nhwc_vs_oracle or (nms_bit_exact_vs_reference_golden and not 34k) or nms_edge_cases or (tile_merge_nms_vs_oracle and (65 or 1000)) or letterbox_kernel_vs_oracle or (model_vs_reference_golden and n_repvgg_128 and plan)'
for t in $tools; do
  log=gpurun_out/${tag}_sanitize_${t}.log
  extra=""
  [ $t = memcheck ] && extra="--leak-check no"
  [ $t = racecheck ] && extra="--racecheck-report all"
  timeout ${SAN_TIMEOUT:-1500} compute-sanitizer --tool $t $extra --error-exitcode 0 --print-limit 30 \
      python -m pytest tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "$SEL" > $log 2>&1
  echo "$t exit $?: $(grep -E 'passed|failed|error' $log | tail -1) | $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1)"
done
