#!/bin/bash
# 2-GPU data-parallel check: bench under torchrun + DP-vs-single-process parity
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/dp2.log 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 \
  scripts/dp_parity.py >> gpurun_out/dp2.log 2>&1; echo "dp_parity exit $?" >> gpurun_out/dp2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
  bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_dp2.json 2>> gpurun_out/dp2.log; echo "bench2 exit $?" >> gpurun_out/dp2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 \
  bench.py --gpus 2 --steps 10 --warmup 3 --batch 8 --extra-batch 0 > gpurun_out/bench_dp2_b8.json 2>> gpurun_out/dp2.log; echo "bench2 b8 exit $?" >> gpurun_out/dp2.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8 --extra-batch 0 --no-cpu-baseline > gpurun_out/bench_dp1_b8.json 2>> gpurun_out/dp2.log
tail -30 gpurun_out/dp2.log; cat gpurun_out/bench_dp2.json gpurun_out/bench_dp2_b8.json gpurun_out/bench_dp1_b8.json | cut -c1-900
