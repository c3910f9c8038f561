#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/phase_prof.log; : > $LOG
for args in "res 1" "res 8" "down2 1" "dconv3 1" "up 1" "head 1" "stem 1" "dconv0 1" "dconv4 1" "down 1" "dconv1 1" "up2 1"; do
  echo "=== $args" >> $LOG
  CGB_PROF=1 timeout 120 unpaired_image_generation_b200/csrc/build/selftest_conv $args >> $LOG 2>&1
  echo "exit $?" >> $LOG
done
for bn in 128 256; do echo "=== res 1 BN=$bn" >> $LOG; CGB_PROF=1 CGB_FORCE_BN=$bn timeout 120 unpaired_image_generation_b200/csrc/build/selftest_conv res 1 >> $LOG 2>&1; done
for bn in 64 128; do echo "=== res 8 BN=$bn" >> $LOG; CGB_PROF=1 CGB_FORCE_BN=$bn timeout 120 unpaired_image_generation_b200/csrc/build/selftest_conv res 8 >> $LOG 2>&1; done
grep -E "===|us/launch|phases|FAIL|PASSED|EXCEPTION|exit [1-9]" $LOG
