#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest5.log
: > $LOG
echo "=== parity res_small" >> $LOG; CGB_PASSES=3 timeout 120 $BIN res_small 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== parity head" >> $LOG; CGB_PASSES=3 timeout 120 $BIN head 1 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== parity res N=8 fprop (persistent: 256 items)" >> $LOG; CGB_PASSES=1 timeout 300 $BIN res 8 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== parity res N=4 dgrad (persistent: 180 items)" >> $LOG; CGB_PASSES=2 timeout 300 $BIN res 4 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== parity dconv3 N=8" >> $LOG; CGB_PASSES=3 timeout 300 $BIN dconv3 8 >> $LOG 2>&1; echo "exit $?" >> $LOG
export CGB_TIMING_ONLY=1 CGB_PROF=1 CGB_PASSES=3
for cfg in "1 0 0" "8 0 0" "8 256 1" "8 128 2" "8 128 1"; do
  set -- $cfg
  echo "=== timing N=$1 BN=$2 MT=$3" >> $LOG
  if [ "$2" = "0" ]; then timeout 120 $BIN res $1 >> $LOG 2>&1; else CGB_FORCE_BN=$2 CGB_FORCE_MT=$3 timeout 120 $BIN res $1 >> $LOG 2>&1; fi
  echo "exit $?" >> $LOG
done
for n in 1 8; do echo "=== head N=$n" >> $LOG; timeout 120 $BIN head $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG; done
grep -E "^===|exit|us/launch|phases|OK|FAIL|patch=" $LOG | cut -c1-260
