#!/bin/bash
# engine check at 256, bench, then the ncu launch list of the same bench command
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python scripts/gpu_engine_check.py 256 > gpurun_out/engine256.log 2>&1; echo "exit $?" >> gpurun_out/engine256.log
grep -v "bias " gpurun_out/engine256.log | grep -vE "res\.[1-7]" | tail -60
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "bench exit $?"
cat gpurun_out/bench_b1.json; tail -5 gpurun_out/bench_b1.err
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu.log; wc -l gpurun_out/launches.csv
