#!/bin/bash
# One GPU-box pass: parity tests, bench (strict fp32 / TF32 / NHWC variants), ncu launch list of the timed region.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh <tag>
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_$TAG.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
cat gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_reference.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 10 --warmup 3 --tf32 1 --no-cpu-baseline --extras 0 --sfod-step 0 > gpurun_out/bench_${TAG}_tf32.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --steps 10 --warmup 3 --channels-last 1 --tf32 1 --no-cpu-baseline --extras 0 --sfod-step 0 > gpurun_out/bench_${TAG}_nhwc_tf32.json 2>> gpurun_out/bench_$TAG.err
timeout 900 python bench.py --workload r101 --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_r101.json 2>> gpurun_out/bench_$TAG.err
timeout 600 python bench.py --workload r101 --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_r101_reference.json 2>> gpurun_out/bench_$TAG.err
timeout 900 python tools/microbench.py > gpurun_out/microbench_$TAG.jsonl 2>> gpurun_out/bench_$TAG.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --sfod-step 0 --profiler-range > gpurun_out/ncu_bench_$TAG.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/bench_$TAG.err
