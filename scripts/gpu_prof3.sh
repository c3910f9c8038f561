#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
LOG=gpurun_out/phase_prof3.log; : > $LOG
for bn in 64 128 256; do for n in 1 8; do echo "=== res $n BN=$bn" >> $LOG; CGB_PROF=1 CGB_FORCE_BN=$bn timeout 120 unpaired_image_generation_b200/csrc/build/selftest_conv res $n >> $LOG 2>&1; done; done
grep -E "===|fprop: [0-9]|dgrad: [0-9]|wgrad: [0-9]|phases|detail" $LOG
