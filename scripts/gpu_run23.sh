set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_cfg4.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=600 -x -k "bvh or cfg4 or mesh or gate" 2>&1 | tail -8
