#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in 0.85 0.5 0.3 0.2; do
  CGB_CTA_FRAC=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extra-batch 0 > gpurun_out/bench_frac_$f.json 2> gpurun_out/bench_frac_$f.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_frac_$f.json') if l.startswith('{')][-1])
r=d['roofline']
print('frac $f: b1 ms', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), '| igemm ms', round(r['ms_per_step'],2), '| wgrad ms', round(r['other_kernels']['wgrad_kernel(tcgen05)']['ms_per_step'],2), '| small', round(r['other_kernels']['wgrad_direct(3-channel layers)']['ms_per_step'],2), '| pw ms', round(r['other_kernels']['instnorm_pointwise']['ms_per_step'],2), r['segments_ms'])
PY
done
