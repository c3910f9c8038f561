#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/phase_prof2.log; : > $LOG
for args in "res 1" "res 8" "up2 1"; do
  echo "=== $args" >> $LOG
  CGB_PROF=1 timeout 120 unpaired_image_generation_b200/csrc/build/selftest_conv $args >> $LOG 2>&1
done
grep -E "===|fprop:|phases|detail" $LOG
