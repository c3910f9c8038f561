#!/bin/bash
# Data-parallel A/B on N GPUs: one short bench line per setting.  N=8 SWEEP='CGB_DP_BUCKETS=2;NCCL_MAX_CTAS=4' bash scripts/gpu_dp_sweep.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}; TAG=${TAG:-r02}
OUT=gpurun_out/${TAG}_dp${N}_sweep.txt
: > $OUT
IFS=';' read -ra SETTINGS <<< "${SWEEP:-CGB_DP_OVERLAP=1}"
port=29540
for s in "${SETTINGS[@]}"; do
  port=$((port+1))
  line=$(env $s timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
         bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-extra-configs 2>/tmp/dp_err.txt)
  python - "$s" "$line" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    print(f"{sys.argv[1]:50s} ms/step {d['ms_per_step']:.3f}  img/s {d['value']:.1f}")
except Exception as ex:
    print(f"{sys.argv[1]:50s} FAILED {ex} {sys.argv[2][:200]}")
PY
done
cat $OUT
