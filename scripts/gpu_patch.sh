#!/bin/bash
# Patch-resident igemm (conv_patch.cu): which UMMA descriptor addressing mode is correct on the hardware, parity
# of every eligible layer shape against the CPU loops, and timing sweeps over (BN, MT).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
LOG=gpurun_out/patch_selftest.log
: > $LOG
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $LOG 2>&1
GOOD=""
for mode in 1 2 4 3; do
  echo "=== mode $mode res_small" >> $LOG
  CGB_PATCH_MODE=$mode CGB_PASSES=3 timeout 120 $BIN res_small 2 >> $LOG 2>&1
  rc=$?
  echo "exit $rc" >> $LOG
  if [ $rc -eq 0 ] && [ -z "$GOOD" ]; then GOOD=$mode; fi
done
echo "GOOD MODE: $GOOD" | tee -a $LOG
if [ -n "$GOOD" ]; then
  export CGB_PATCH_MODE=$GOOD
  for c in res head stem dconv3 dconv4 down up; do
    echo "=== mode $GOOD $c" >> $LOG
    CGB_PASSES=3 timeout 300 $BIN $c 1 >> $LOG 2>&1
    echo "exit $?" >> $LOG
  done
  echo "=== mode $GOOD res N=2 parity (MT=2 forced)" >> $LOG
  CGB_FORCE_MT=2 CGB_FORCE_BN=256 CGB_PASSES=3 timeout 300 $BIN res 2 >> $LOG 2>&1; echo "exit $?" >> $LOG
  export CGB_TIMING_ONLY=1
  for n in 1 8; do
    for bn in 64 128 256; do for mt in 1 2; do
      echo "=== timing res N=$n BN=$bn MT=$mt" >> $LOG
      CGB_FORCE_BN=$bn CGB_FORCE_MT=$mt CGB_PASSES=3 timeout 120 $BIN res $n >> $LOG 2>&1; echo "exit $?" >> $LOG
    done; done
    echo "=== timing res N=$n old kernel" >> $LOG
    CGB_PATCH_MODE=0 CGB_PASSES=3 timeout 120 $BIN res $n >> $LOG 2>&1; echo "exit $?" >> $LOG
    echo "=== timing res N=$n default choice" >> $LOG
    CGB_PASSES=3 timeout 120 $BIN res $n >> $LOG 2>&1; echo "exit $?" >> $LOG
  done
fi
grep -E "^===|^case|exit|OK|FAIL|us/launch|EXCEPTION|timeout|rror|GOOD|patch=" $LOG | tail -150
if [ -n "$GOOD" ]; then
  for n in 1 8; do
    for c in head stem; do
      echo "=== timing $c 256x256 N=$n patch" >> $LOG
      CGB_PASSES=3 timeout 120 $BIN $c $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG
      echo "=== timing $c 256x256 N=$n old" >> $LOG
      CGB_PATCH_MODE=0 CGB_PASSES=3 timeout 120 $BIN $c $n 0 256 >> $LOG 2>&1; echo "exit $?" >> $LOG
    done
  done
  grep -E "^===|us/launch|patch=" $LOG | tail -40
fi
