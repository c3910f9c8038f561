#!/bin/bash
# ncu evidence: (1) full capture of the dominant kernel on the residual-block shape (batch 1 and 8),
# (2) launch list of a short bench run.  Each profiled command first exits 0 without ncu.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
BIN=unpaired_image_generation_b200/csrc/build/selftest_conv
$BIN res 8 > gpurun_out/plain_res8.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:igemm_conv -s 2 -c 2 -o gpurun_out/prof_igemm_res8 \
  $BIN res 8 > gpurun_out/ncu_res8.log 2>&1
echo "ncu res8 exit $?"
$BIN res 1 > gpurun_out/plain_res1.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"igemm_conv|wgrad_kernel" -s 2 -c 3 -o gpurun_out/prof_igemm_res1 \
  $BIN res 1 > gpurun_out/ncu_res1.log 2>&1
echo "ncu res1 exit $?"
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/ncu.log 2>&1
echo "ncu launches exit $?"; wc -l gpurun_out/launches.csv; ls -la gpurun_out/*.ncu-rep
