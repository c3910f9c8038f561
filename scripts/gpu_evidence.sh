#!/bin/bash
# Round evidence in one gpurun call: default bench line + reference arm, per-op tables / timelines, ncu launch lists and
# ncu --set full captures of the dominant kernels.   TAG=r02_final bash scripts/gpu_evidence.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${TAG:-r02_final}
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps ${REF_STEPS:-5} --warmup 2 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; echo "reference exit $?"
TAG=$TAG BATCHES="1 8" NCU_LIST=1 bash scripts/gpu_profile.sh
B="--steps 1 --warmup 3 --no-cpu-baseline --no-extra-configs"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"igemm_patch_kernel|wgrad_pair" -s 200 -c 14 \
  -o gpurun_out/${TAG}_ncu_full_conv_b1 -f python bench.py --batch 1 $B > gpurun_out/${TAG}_ncu_full_conv_b1.log 2>&1; echo "ncu conv b1 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"igemm_patch_kernel<256, 1, 3|wgrad_pair" -s 120 -c 8 \
  -o gpurun_out/${TAG}_ncu_full_conv_b8 -f python bench.py --batch 8 $B > gpurun_out/${TAG}_ncu_full_conv_b8.log 2>&1; echo "ncu conv b8 exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"in_apply_bulk|in_bwd_reduce_bulk|in_bwd_apply_bulk" -s 150 -c 12 \
  -o gpurun_out/${TAG}_ncu_full_in_b8 -f python bench.py --batch 8 $B > gpurun_out/${TAG}_ncu_full_in_b8.log 2>&1; echo "ncu IN b8 exit $?"
