#!/bin/bash
# ncu evidence for the current code: (1) --set full captures of the dominant conv kernels on the residual-block
# shape (batch 8 and batch 1), (2) launch lists of short bench runs at batch 1 and batch 8.
# Every profiled command first exits 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
export CGB_TIMING_ONLY=1
$BIN res 8 > gpurun_out/plain_res8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"igemm_patch|wgrad_kernel" -s 3 -c 3 -o gpurun_out/prof_patch_res8 -f \
  $BIN res 8 > gpurun_out/ncu_res8.log 2>&1
echo "ncu res8 exit $?"
$BIN res 1 > gpurun_out/plain_res1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"igemm_patch|wgrad_kernel" -s 3 -c 3 -o gpurun_out/prof_patch_res1 -f \
  $BIN res 1 > gpurun_out/ncu_res1.log 2>&1
echo "ncu res1 exit $?"
unset CGB_TIMING_ONLY
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/plain_b1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_b1.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/ncu_b1.log 2>&1
echo "ncu launches b1 exit $?"; wc -l gpurun_out/launches_b1.csv
python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/plain_b8.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/launches_b8.csv \
  python bench.py --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/ncu_b8.log 2>&1
echo "ncu launches b8 exit $?"; wc -l gpurun_out/launches_b8.csv
ls -la gpurun_out/*.ncu-rep
