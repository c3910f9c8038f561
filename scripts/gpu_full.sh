#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_selftest.sh > gpurun_out/selftest_summary.txt 2>&1
grep -E "^case|FAIL|EXCEPTION|exit [1-9]" gpurun_out/selftest_summary.txt | grep -v PASSED | head
echo "selftest passed: $(grep -c PASSED gpurun_out/selftest_summary.txt)"
grep -E "us/launch" gpurun_out/selftest_conv.log | head -60 | tr '\n' ';' | cut -c1-2500; echo
CGB_REQUIRE_GRAPH=1 bash scripts/gpu_tests_bench.sh > gpurun_out/tests_bench.txt 2>&1
grep -E "passed|failed|FAILED|Error|bench exit" gpurun_out/tests_bench.txt | head
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench.json') if l.startswith('{')][-1])
r=d['roofline']; e=d.get('extra_batch',{})
print('b1 ms', round(d['ms_per_step'],2), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), '| igemm ms', round(r['ms_per_step'],2), 'TF', round(r['achieved'],1), '| wgrad ms', round(r['other_kernels']['wgrad_kernel(tcgen05)']['ms_per_step'],2), '| small', round(r['other_kernels']['wgrad_direct(3-channel layers)']['ms_per_step'],2),'| pw ms', round(r['other_kernels']['instnorm_pointwise']['ms_per_step'],2), r['segments_ms'])
print('b8', e)
PY
