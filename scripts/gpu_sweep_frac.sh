#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_selftest.sh > gpurun_out/selftest_summary.txt 2>&1
grep -E "^case|FAIL|EXCEPTION|exit [1-9]" gpurun_out/selftest_summary.txt | grep -v PASSED | head
grep -c PASSED gpurun_out/selftest_summary.txt
for f in 0.85 0.5 0.25 0.12; do
  CGB_CTA_FRAC=$f timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extra-batch 8 > gpurun_out/bench_frac_$f.json 2> gpurun_out/bench_frac_$f.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_frac_$f.json') if l.startswith('{')][-1])
r=d['roofline']; e=d.get('extra_batch',{})
print('frac $f: b1 ms', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), '| igemm ms', round(r['ms_per_step'],2), 'TF', round(r['achieved'],1), '| wgrad ms', round(r['other_kernels']['wgrad_kernel(tcgen05)']['ms_per_step'],2), '| pw ms', round(r['other_kernels']['instnorm_pointwise']['ms_per_step'],2), '| b8 ms', round(e.get('ms_per_step',0),2), 'img/s', round(e.get('value',0),1), 'igemm TF', round(e.get('igemm_tflops',0),1))
PY
done
