#!/bin/bash
# ncu --set full capture of the hot-path kernels inside one bench step (1 GPU only). Usage: bash tools/gpu_profile.sh <tag>
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k 'regex:roi_align|bn_apply|bn_stats|bn_finalize|ema_multi|nms_mask|nms_scan|bitonic|rpn_|frcnn_|transpose' \
    -o gpurun_out/prof_${TAG}_step -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --profiler-range > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
