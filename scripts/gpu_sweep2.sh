#!/bin/bash
# two-variable sweep at batch 1: "VAR1=a VAR2=b" combinations listed in COMBOS (semicolon separated)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
IFS=';' read -ra CS <<< "$COMBOS"
for c in "${CS[@]}"; do
  env $c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --extra-batch ${EXTRA:-0} > gpurun_out/bench_combo.json 2> gpurun_out/bench_combo.err
  python - "$c" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/bench_combo.json"))
    r = d["roofline"]
    e = d.get("extra_batch") or {}
    print("%-50s b1 %.3f ms (%.1f img/s) serial igemm %.2f wgrad %.2f pw %.2f | b8 %s" % (sys.argv[1], d["ms_per_step"], d["value"], r["all_igemm_launches"]["ms_per_step_serial"], r["other_kernels"]["wgrad_kernel(tcgen05)"]["ms_per_step"], r["other_kernels"]["instnorm_pointwise"]["ms_per_step"], ("%.2f ms" % e["ms_per_step"]) if e else "-"))
except Exception as ex:
    print(sys.argv[1], "parse error", ex)
PY
done
