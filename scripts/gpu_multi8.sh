#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_multi.sh 8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 --warmup 5 --workload cfg5 --no-side > gpurun_out/bench_cfg5_n8.json 2> gpurun_out/bench_cfg5_n8.err
echo "cfg5 rc=$?"; python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg5_n8.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","kernel_ms_per_step","e2e","frame_check"):
        print(k, json.dumps(d.get(k))[:600])
except Exception as e: print("no json", e)
EOF
