#!/bin/bash
# 2-GPU pass: IPC / NCCL tests again + compute-sanitizer (memcheck, racecheck) over the peer-exchange kernels (local ring on one GPU).
set -u
TAG=${1:-r2v}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "peer or nccl" > gpurun_out/pytest_p2p_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p2p_$TAG.log
tail -4 gpurun_out/pytest_p2p_$TAG.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 0 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "peer_statistic_exchange" \
    > gpurun_out/sanitizer_memcheck_$TAG.log 2>&1; echo "memcheck rc=$?"
grep -E "passed|failed|ERROR SUMMARY|Invalid|hit CUDA" gpurun_out/sanitizer_memcheck_$TAG.log | head -10
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 0 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "peer_statistic_exchange" \
    > gpurun_out/sanitizer_racecheck_$TAG.log 2>&1; echo "racecheck rc=$?"
grep -E "passed|failed|RACECHECK SUMMARY|hazard" gpurun_out/sanitizer_racecheck_$TAG.log | head -10
