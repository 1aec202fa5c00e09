set -u
mkdir -p gpurun_out
timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
timeout 300 python scripts/gpu_stride.py 2>&1 | tail -5
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=300 2>&1 | tail -5
