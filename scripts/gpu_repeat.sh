#!/bin/bash
# Repeat one bench configuration N times and keep stdout/stderr of failing runs (flakiness hunt).
#   TAG=r02_z N=6 ENVS="CGB_WGRADP_MIN_CHUNKS=32" BATCH=1 bash scripts/gpu_repeat.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02}
OUT=gpurun_out/${TAG}_repeat.txt
: > $OUT
for i in $(seq 1 ${N:-5}); do
  env ${ENVS:-CGB_PDL=1} timeout 200 python bench.py --batch ${BATCH:-1} --steps ${STEPS:-20} --no-cpu-baseline --no-extra-configs > /tmp/rep_out.txt 2> /tmp/rep_err.txt
  rc=$?
  ms=$(python -c "import json;print(json.load(open('/tmp/rep_out.txt'))['ms_per_step'])" 2>/dev/null)
  echo "run $i rc=$rc ms_per_step=$ms" >> $OUT
  if [ "$rc" != "0" ] || [ -z "$ms" ]; then
    echo "---- stdout" >> $OUT; head -c 2000 /tmp/rep_out.txt >> $OUT
    echo "---- stderr" >> $OUT; tail -n 40 /tmp/rep_err.txt >> $OUT
  fi
done
cat $OUT | cut -c1-400
