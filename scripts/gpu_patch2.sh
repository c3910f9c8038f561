#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest2.log
: > $LOG
export CGB_TIMING_ONLY=1 CGB_PROF=1 CGB_PASSES=1
for mode in 1 3 5 0; do
  for cfg in "1 64 1" "8 256 1" "8 128 1"; do
    set -- $cfg
    echo "=== mode $mode N=$1 BN=$2 MT=$3" >> $LOG
    CGB_PATCH_MODE=$mode CGB_FORCE_BN=$2 CGB_FORCE_MT=$3 timeout 120 $BIN res $1 >> $LOG 2>&1; echo "exit $?" >> $LOG
  done
done
echo "=== mode 5 parity" >> $LOG
CGB_TIMING_ONLY= CGB_PATCH_MODE=5 env -u CGB_TIMING_ONLY timeout 120 $BIN res_small 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
grep -E "^===|exit|us/launch|phases|detail|patch=|OK|FAIL" $LOG
