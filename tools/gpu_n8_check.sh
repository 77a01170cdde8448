#!/bin/bash
# 8-GPU pass: bench (headline + AdaBN record with the peer-memory exchange and the NCCL baseline).
set -u
TAG=${1:-r2s}
N=${2:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench rc=$?"
python - $TAG $N <<'PY'
import json,sys
try:
    d=json.loads(open("gpurun_out/bench_%s_n%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print(json.dumps({k:d.get(k) for k in ("value","n_gpus","ms_per_step","e2e","adabn","sfod_step")}, indent=1))
except Exception as e:
    print("parse failed", e)
PY
tail -5 gpurun_out/bench_${TAG}_n$N.err
