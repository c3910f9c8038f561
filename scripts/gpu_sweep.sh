#!/bin/bash
# A/B sweeps of the library's experiment switches: one short bench run per setting.
#   TAG=r02_d SWEEP='CGB_WGRAD_LANES=1;CGB_WGRAD_LANES=2;CGB_WGRAD_LANES=2 CGB_WGRAD_MIN_CHUNKS=32' [BATCH=1] bash scripts/gpu_sweep.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
OUT=gpurun_out/${TAG}_sweep_b${BATCH:-1}.txt
: > $OUT
IFS=';' read -ra SETTINGS <<< "${SWEEP:-CGB_PDL=1}"
for s in "${SETTINGS[@]}"; do
  line=$(env $s timeout 300 python bench.py --batch ${BATCH:-1} --size ${SIZE:-256} --steps ${STEPS:-20} --no-cpu-baseline --no-extra-configs 2>/tmp/sweep_err.txt)
  [ -z "$line" ] && { echo "---- $s: no output; stderr tail:" >> $OUT; tail -n 15 /tmp/sweep_err.txt >> $OUT; }
  python - "$s" "$line" >> $OUT <<'PY'
import json, sys
try:
    d = json.loads(sys.argv[2])
    r = d["roofline"]
    print(f"{sys.argv[1]:60s} ms/step {d['ms_per_step']:.3f}  img/s {d['value']:.1f}  res fprop/dgrad/wgrad TF/s "
          f"{r['res_block_conv_tflops']}  pointwise serial ms {r['other_kernels']['instnorm_pointwise']['ms_per_step']:.2f}")
except Exception as ex:
    print(f"{sys.argv[1]:60s} FAILED {ex} {sys.argv[2][:200]}")
PY
done
cat $OUT
