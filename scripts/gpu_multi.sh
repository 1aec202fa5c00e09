#!/bin/bash
# bench.py on N GPUs of one box (gpurun --gpus N): the driver's launch line
N=${1:-2}; shift || true
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 "$@" > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "rc=$?"; tail -5 gpurun_out/bench_n$N.err; python - <<EOF
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
    for k in ("value","ms_per_step","kernel_ms_per_step","per_rank_kernel_step_draw_ms","e2e","frame_check","also_4k","nccl_row_tiles","gpu_launches"):
        print(k, json.dumps(d.get(k))[:700])
except Exception as e: print("no json", e)
EOF
