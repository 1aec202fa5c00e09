set -u
echo "== run-2 commit (673335b)"; (cd _wt_run2 && timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py; timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4)
echo "== current"; timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py; timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
echo "== current, RT_COOP_RAYS=0"; UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_nocoop.so timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=300 2>&1 | tail -6
