set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_cfg4_bands.py 2>&1 | tail -5
