#!/bin/bash
# compute-sanitizer over the parity tests of the hand-written kernels (runs on the GPU box via gpurun).
#   tools/gpu_sanitize.sh <tag> [pytest -k expression]      -> gpurun_out/sanitize_<tag>_{memcheck,racecheck,synccheck}.txt (+ .summary)
set -u
tag=$1
OUT=gpurun_out
mkdir -p $OUT
SEL=${2:-'irblock or dwproj or stem_kernel or conv2d_against or cta_pair or conv_chain or decoder_parity or decoder_integer or combined_nms or loss_forward or hard_negative or iou or match or prior or cuda_reproduces or cuda_priors or cuda_losses or cuda_decoder or cuda_augmentation'}
for tool in memcheck racecheck synccheck; do
  log=$OUT/sanitize_${tag}_${tool}.txt
  timeout ${SAN_TIMEOUT:-700} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 0 \
      python -m pytest tests/test_net_gpu.py tests/test_box_gpu.py tests/test_nms_gpu.py tests/test_loss_gpu.py tests/test_golden.py tests/test_ref_golden.py \
      -m gpu -q -x -k "$SEL" -p no:cacheprovider > $log 2>&1
  echo "exit=$?" >> $log
  { echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit=" $log | tail -5; grep -cE "^=========     (Invalid|Race|Barrier|Uninit|Hazard)" $log; } > $OUT/sanitize_${tag}_${tool}.summary
  cat $OUT/sanitize_${tag}_${tool}.summary
  # keep only the head / tail of very long reports
  if [ $(wc -c < $log) -gt 400000 ]; then head -c 200000 $log > $log.tmp; echo "... [truncated] ..." >> $log.tmp; tail -c 100000 $log >> $log.tmp; mv $log.tmp $log; fi
done
