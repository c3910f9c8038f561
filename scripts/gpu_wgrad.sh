#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/wgrad_selftest.log
: > $LOG
for c in res_small down up dconv1 dconv3; do
  echo "=== parity $c" >> $LOG; CGB_PASSES=4 timeout 200 $BIN $c 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
done
echo "=== parity res" >> $LOG; CGB_PASSES=4 timeout 200 $BIN res 1 >> $LOG 2>&1; echo "exit $?" >> $LOG
export CGB_TIMING_ONLY=1 CGB_PASSES=4
for n in 1 8; do echo "=== timing res N=$n" >> $LOG; timeout 120 $BIN res $n >> $LOG 2>&1; echo "exit $?" >> $LOG; done
grep -E "^===|exit|us/launch|OK|FAIL|split_k|EXCEPTION|rror" $LOG | cut -c1-200
