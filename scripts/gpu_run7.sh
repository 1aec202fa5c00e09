set -u
mkdir -p gpurun_out
timeout 600 bash scripts/gpu_variants.sh head cfg2 cfg3 2>&1 | tee gpurun_out/variants.log
timeout 300 python scripts/gpu_stride.py 2>&1 | tail -5
UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_m2.so timeout 300 python scripts/gpu_stride.py 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "benchmarked or unsplit or split_lane" --timeout=300 2>&1 | tail -5
