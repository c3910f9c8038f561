#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/stage_sweep.log; : > $LOG
for bin in build build_deep; do
  for bn in 64 128 256; do
    for n in 1 8; do
      echo "=== $bin BN=$bn N=$n" >> $LOG
      CGB_FORCE_BN=$bn timeout 120 unpaired_image_generation_b200/csrc/$bin/selftest_conv res $n >> $LOG 2>&1
    done
  done
done
grep -E "===|fprop:|dgrad:|wgrad:|FAIL" $LOG
