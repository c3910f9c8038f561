#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest4.log
: > $LOG
echo "=== parity" >> $LOG
CGB_PASSES=3 timeout 120 $BIN res_small 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
CGB_PASSES=3 timeout 120 $BIN head 1 >> $LOG 2>&1; echo "exit $?" >> $LOG
export CGB_TIMING_ONLY=1 CGB_PROF=1 CGB_PASSES=${PASSES:-1}
for cfg in "1 64 1" "1 128 1" "8 256 1" "8 256 2" "8 128 2" "8 128 1"; do
  set -- $cfg
  echo "=== N=$1 BN=$2 MT=$3" >> $LOG
  CGB_FORCE_BN=$2 CGB_FORCE_MT=$3 timeout 120 $BIN res $1 >> $LOG 2>&1; echo "exit $?" >> $LOG
done
for n in 1 8; do
  echo "=== head N=$n" >> $LOG
  timeout 120 $BIN head $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG
done
grep -E "^===|exit|us/launch|phases|OK|FAIL" $LOG
