#!/bin/bash
# compute-sanitizer over the hand-written kernels (SURVEY section 4.2 kernel tier).  Output (not just exit codes) is kept:
# round 1 lost the reason for its "exit 86" because the tool's text was discarded.
#   TAG=r02 bash scripts/gpu_sanitize.sh        -> gpurun_out/${TAG}_sanitizer.txt
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
OUT=gpurun_out/${TAG}_sanitizer.txt
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
: > $OUT
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  echo "=== $tool: $name" >> $OUT
  CGB_NO_GRAPH=1 CGB_PDL=${SAN_PDL:-1} timeout ${SAN_TIMEOUT:-600} compute-sanitizer --tool $tool --launch-timeout 120 \
    --print-limit 20 --error-exitcode 77 "$@" > gpurun_out/${TAG}_san_tmp.log 2>&1
  local rc=$?
  grep -E "^=========|ERROR SUMMARY|error|Error|hazard|exit|OK|ok|passed|failed" gpurun_out/${TAG}_san_tmp.log | head -n 40 >> $OUT
  echo "rc $rc" >> $OUT
  echo "$tool $name rc $rc"
}
compute-sanitizer --version >> $OUT 2>&1
# 0. does the tool attach at all on this box?  (a trivial program, no tcgen05)
cat > /tmp/san_probe.cu <<'EOC'
#include <cstdio>
__global__ void k(int* p) { p[threadIdx.x] = threadIdx.x; }
int main() { int* p; cudaMalloc(&p, 128); k<<<1, 32>>>(p); printf("probe %s\n", cudaGetErrorString(cudaDeviceSynchronize())); return 0; }
EOC
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/san_probe /tmp/san_probe.cu >> $OUT 2>&1
run probe memcheck /tmp/san_probe
# 1. the conv kernels through the stand-alone self-test (tcgen05 / TMA / mbarrier)
for c in ${SAN_CASES:-res_small head stem down up dconv4}; do
  run $c memcheck $BIN $c 1
done
for c in ${SAN_SYNC_CASES:-res_small down up}; do
  run $c synccheck $BIN $c 1
  run $c racecheck $BIN $c 1
done
# 2. a whole 64x64 train step through the C ABI (every kernel of the product path, eager launches)
run "train step 64x64 (bf16)" memcheck python scripts/san_step.py bf16
run "train step 64x64 (fp32 validation mode)" memcheck python scripts/san_step.py fp32
rm -f gpurun_out/${TAG}_san_tmp.log
tail -n 60 $OUT
