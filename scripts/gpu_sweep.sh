#!/bin/bash
# sweep one environment variable over the full-step bench: VAR=<name> VALS="a b c"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in $VALS; do
  env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/bench_${VAR}_$v.json 2> gpurun_out/bench_${VAR}_$v.err
  echo "$VAR=$v exit $?"
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${VAR}_$v.json"))
    r = d["roofline"]
    print("  b1 ms/step %.3f img/s %.1f igemm %.3f ms wgrad %.3f ms pointwise %.3f ms seg %s" % (d["ms_per_step"], d["value"], r["all_igemm_launches"]["ms_per_step_serial"], r["other_kernels"]["wgrad_kernel(tcgen05)"]["ms_per_step"], r["other_kernels"]["instnorm_pointwise"]["ms_per_step"], {k: round(x, 3) for k, x in r["segments_ms"].items()}))
    e = d.get("extra_batch")
    if e: print("  b8 ms/step %.3f img/s %.1f igemm TF %.1f wgrad TF %.1f res %s" % (e["ms_per_step"], e["value"], e["igemm_tflops"], e["wgrad_tflops"], e.get("res_block_conv_tflops")))
except Exception as ex:
    print("  parse error", ex)
PY
  tail -3 gpurun_out/bench_${VAR}_$v.err
done
