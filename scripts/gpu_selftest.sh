#!/bin/bash
# Stand-alone conv self-test (CPU loop nests as the reference) over a list of settings.
#   TAG=r02_m CASES="res_small:1 res:1 res:8 head:1:256" SETTINGS="CGB_PATCH_CG=0;CGB_PATCH_CG=1" bash scripts/gpu_selftest.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
OUT=gpurun_out/${TAG}_selftest.txt
: > $OUT
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
IFS=';' read -ra SETS <<< "${SETTINGS:-CGB_PDL=1}"
for s in "${SETS[@]}"; do
  for c in ${CASES:-res_small:1 res:1}; do
    IFS=':' read -r name n h <<< "$c"   # name:batch[:spatial extent]
    echo "=== $s $name N=$n H=${h:-default}" >> $OUT
    env $s timeout ${CASE_TIMEOUT:-120} $BIN $name $n 0 ${h:-0} >> $OUT 2>&1
    echo "exit $?" >> $OUT
  done
done
grep -E "===|us/launch|PASSED|FAILED|exit [1-9]|timeout|BN=" $OUT | cut -c1-220
