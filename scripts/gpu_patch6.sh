#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest6.log
: > $LOG
for c in head stem dconv4 res_small; do
  echo "=== parity $c" >> $LOG; CGB_PASSES=3 timeout 120 $BIN $c 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
done
export CGB_TIMING_ONLY=1 CGB_PASSES=3
for n in 1 8; do for c in head stem; do
  echo "=== timing $c 256 N=$n" >> $LOG; timeout 120 $BIN $c $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG
  echo "=== timing $c 256 N=$n (no 16-channel patches)" >> $LOG; CGB_PATCH_NO16=1 timeout 120 $BIN $c $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG
done; done
grep -E "^===|exit|us/launch|OK|FAIL|patch=|EXCEPTION|rror" $LOG | cut -c1-200
