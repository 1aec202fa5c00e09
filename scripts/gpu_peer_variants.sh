#!/bin/bash
# scripts/gpu_peer_cost.py under torchrun for the default library and every kernel-variant library (N = $1 ranks)
N=${1:-2}; shift
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/gpu_peer_cost.py "$@" 2>&1 | grep "step "; }
echo "== default"; run "$@"
for f in uob_raytracer_b200/variants/var_*.so; do echo "== $f"; UOB_RT_LIB=$PWD/$f run "$@"; done
