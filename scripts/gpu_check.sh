#!/bin/bash
# One gpurun call = GPU test-suite + bench lines.  Usage (from the repo root, through gpurun):
#   TAG=r02_a [TESTS="tests -m gpu"] [BENCH=1] [REF=1] [SANITIZE=1] bash scripts/gpu_check.sh
# Everything lands in gpurun_out/${TAG}_*.  Each stage is wrapped in its own timeout so a hang cannot hold the box.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
TESTS=${TESTS:-tests -m gpu}
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout ${TEST_TIMEOUT:-1500} python -m pytest $TESTS -x -q -s > gpurun_out/${TAG}_pytest_gpu.log 2>&1
  echo "pytest exit $?"; tail -n 25 gpurun_out/${TAG}_pytest_gpu.log
fi
if [ "${BENCH:-1}" = "1" ]; then
  timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
  echo "bench exit $?"; cut -c1-1500 gpurun_out/${TAG}_bench.json; tail -n 5 gpurun_out/${TAG}_bench.err
fi
if [ "${REF:-0}" = "1" ]; then
  timeout 600 python bench.py --impl reference --steps ${REF_STEPS:-5} --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
  echo "bench reference exit $?"; cut -c1-600 gpurun_out/${TAG}_bench_reference.json
fi
if [ "${SANITIZE:-0}" = "1" ]; then
  bash scripts/gpu_sanitize.sh
fi
