set -u
mkdir -p gpurun_out
timeout 600 bash scripts/gpu_variants.sh head cfg2 cfg3 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
timeout 300 python scripts/gpu_stride.py > gpurun_out/stride_default.log 2>&1; cat gpurun_out/stride_default.log
timeout 900 python -m pytest tests -q -m gpu -rA --timeout=600 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -30
