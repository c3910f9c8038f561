#!/bin/bash
# Round evidence: GPU tests, the default bench line, ncu launch lists (batch 1 and 8) and --set full captures of the
# dominant kernels.  Every profiled command first exits 0 without ncu.  TAG names the outputs (e.g. r01_e).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r01_e}
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "bench reference exit $?"
CGB_PROFILE_OPS=1 python scripts/gpu_timeline.py 1 > gpurun_out/${TAG}_ops_b1.txt 2>&1
CGB_PROFILE_OPS=1 python scripts/gpu_timeline.py 8 > gpurun_out/${TAG}_ops_b8.txt 2>&1
python scripts/gpu_timeline.py 1 > gpurun_out/${TAG}_timeline_b1.txt 2>&1
python scripts/gpu_timeline.py 8 > gpurun_out/${TAG}_timeline_b8.txt 2>&1
for n in 8 1; do
  CGB_TIMING_ONLY=1 $BIN res $n > gpurun_out/${TAG}_plain_res$n.log 2>&1 &&
  CGB_TIMING_ONLY=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:"igemm_patch|wgrad_kernel" -s 3 -c 3 \
    -o gpurun_out/${TAG}_ncu_res$n -f $BIN res $n > gpurun_out/${TAG}_ncu_res$n.log 2>&1
  echo "ncu res$n exit $?"
done
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/${TAG}_plain_b1.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/${TAG}_launches_b1.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/${TAG}_ncu_b1.log 2>&1
echo "ncu launches b1 exit $?"; wc -l gpurun_out/${TAG}_launches_b1.csv
python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/${TAG}_plain_b8.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/${TAG}_launches_b8.csv \
  python bench.py --batch 8 --steps 1 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/${TAG}_ncu_b8.log 2>&1
echo "ncu launches b8 exit $?"; wc -l gpurun_out/${TAG}_launches_b8.csv
cat gpurun_out/${TAG}_bench.json | cut -c1-600
