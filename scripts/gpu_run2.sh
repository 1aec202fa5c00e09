set -u
mkdir -p gpurun_out
bash scripts/gpu_variants.sh head cfg2 cfg3 > gpurun_out/variants.log 2>&1; cat gpurun_out/variants.log
python scripts/gpu_stride.py > gpurun_out/stride_default.log 2>&1; cat gpurun_out/stride_default.log
UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_base.so python scripts/gpu_stride.py > gpurun_out/stride_base.log 2>&1; cat gpurun_out/stride_base.log
python -m pytest tests -q -m gpu -rA --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -30
python tests/tools/fuzz_diag.py > gpurun_out/fuzz_diag.log 2>&1; cat gpurun_out/fuzz_diag.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
