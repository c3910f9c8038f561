#!/bin/bash
# 2-GPU data parallel: parity (2 ranks x batch 1 == batch 2), merged vs segmented DP step at batch 1, batch 8
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29611 scripts/dp_parity.py > gpurun_out/dp2.log 2>&1; echo "dp_parity exit $?"; grep -E "DP parity|identical" gpurun_out/dp2.log | cut -c1-200
run 29612 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/bench_dp2.json 2>> gpurun_out/dp2.log; echo "bench2 merged exit $?"
CGB_DP_SEGMENTED=1 run 29613 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/bench_dp2_seg.json 2>> gpurun_out/dp2.log; echo "bench2 segmented exit $?"
run 29614 bench.py --gpus 2 --steps 10 --warmup 3 --batch 8 --no-cpu-baseline --extra-batch 0 > gpurun_out/bench_dp2_b8.json 2>> gpurun_out/dp2.log; echo "bench2 b8 exit $?"
for f in bench_dp2 bench_dp2_seg bench_dp2_b8; do python -c "
import json; d=json.load(open('gpurun_out/$f.json')); print('$f', round(d['ms_per_step'],3), 'ms', round(d['value'],1), 'img/s e2e', round(d['e2e']['value'],1))"; done
tail -3 gpurun_out/dp2.log
