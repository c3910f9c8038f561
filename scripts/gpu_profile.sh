#!/bin/bash
# Warm per-op timings + lane timeline of the step (CUDA events, no profiler), optional ncu launch list / full capture.
#   TAG=r02_b [BATCHES="1 8"] [NCU_LIST=1] [NCU_FULL="regex"] bash scripts/gpu_profile.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
for b in ${BATCHES:-1 8}; do
  CGB_PROFILE_OPS=1 timeout 300 python scripts/gpu_timeline.py $b > gpurun_out/${TAG}_ops_b$b.txt 2>&1
  timeout 300 python scripts/gpu_timeline.py $b > gpurun_out/${TAG}_timeline_b$b.txt 2>&1
  echo "ops/timeline b$b exit $?"
done
if [ "${NCU_LIST:-0}" = "1" ]; then
  for b in ${BATCHES:-1 8}; do
    timeout 300 python bench.py --batch $b --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/${TAG}_plain_b$b.log 2>&1 &&
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/${TAG}_launches_b$b.csv \
      python bench.py --batch $b --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/${TAG}_ncu_b$b.log 2>&1
    echo "ncu launch list b$b exit $?"
    python scripts/launch_summary.py gpurun_out/${TAG}_launches_b$b.csv > gpurun_out/${TAG}_launches_b${b}_summary.txt 2>&1
  done
fi
if [ -n "${NCU_FULL:-}" ]; then
  for b in ${BATCHES:-1 8}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"${NCU_FULL}" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-4} \
      -o gpurun_out/${TAG}_ncu_full_b$b -f python bench.py --batch $b --steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs \
      > gpurun_out/${TAG}_ncu_full_b$b.log 2>&1
    echo "ncu full b$b exit $?"
  done
fi
