#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest3.log
: > $LOG
echo "=== parity rot=1 bsplit=2" >> $LOG
CGB_PATCH_ROT=1 CGB_PATCH_BSPLIT=2 CGB_PASSES=3 timeout 120 $BIN res_small 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
CGB_PATCH_ROT=1 CGB_PATCH_BSPLIT=2 CGB_PASSES=3 timeout 120 $BIN res 1 >> $LOG 2>&1; echo "exit $?" >> $LOG
export CGB_TIMING_ONLY=1 CGB_PROF=1 CGB_PASSES=1
for rot in 0 1; do for sp in 1 2 4; do
  for cfg in "1 64 1" "8 256 1" "8 256 2" "8 128 2"; do
    set -- $cfg
    echo "=== rot $rot bsplit $sp N=$1 BN=$2 MT=$3" >> $LOG
    CGB_PATCH_ROT=$rot CGB_PATCH_BSPLIT=$sp CGB_FORCE_BN=$2 CGB_FORCE_MT=$3 timeout 120 $BIN res $1 >> $LOG 2>&1; echo "exit $?" >> $LOG
  done
done; done
grep -E "^===|exit|us/launch|phases|OK|FAIL" $LOG
