#!/bin/bash
# Runs every case of the stand-alone conv self-test, each in its own process under a timeout
# so a protocol bug (trap / hang) in one case cannot take the others down.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/selftest_conv.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
for c in ${CASES:-res_small res down down2 dconv1 dconv3 up up2 stem head dconv0 dconv4}; do
  echo "=== $c" >> $LOG
  timeout 120 $BIN $c ${NBATCH:-1} >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
echo "=== res no-cluster" >> $LOG; CGB_NO_CLUSTER=1 timeout 120 $BIN res 1 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== res N=8 no-cluster" >> $LOG; CGB_NO_CLUSTER=1 timeout 120 $BIN res 8 >> $LOG 2>&1; echo "exit $?" >> $LOG
echo "=== res N=8" >> $LOG; timeout 120 $BIN res 8 >> $LOG 2>&1; echo "exit $?" >> $LOG
grep -E "^case|exit|OK|FAIL|us/launch|EXCEPTION|timeout|error" $LOG | tail -120
