#!/bin/bash
# 2-GPU pass for the NVLink peer-memory statistic exchange: tests (local ring on one GPU, NCCL and IPC workers on two), bench at N = 2.
set -u
TAG=${1:-r2p}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$TAG.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_properties.py -m gpu -x -q -k "peer or nccl or bn" > gpurun_out/pytest_p2p_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p2p_$TAG.log
tail -15 gpurun_out/pytest_p2p_$TAG.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 10 --warmup 3 \
    --sfod-step 0 > gpurun_out/bench_${TAG}_n2.json 2> gpurun_out/bench_${TAG}_n2.err; echo "bench rc=$?"
python - $TAG <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/bench_%s_n2.json" % sys.argv[1] if len(sys.argv)>1 else "gpurun_out/bench_r2p_n2.json").read().strip().splitlines()[-1])
    print(json.dumps({k:d[k] for k in ("value","n_gpus","ms_per_step","adabn")}, indent=1))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/bench_${TAG}_n2.err
