set -u
echo "== run-2 commit (673335b)"; (cd _wt_run2 && timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py)
echo "== current"; timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
echo "== current + f32x2"; UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_f32x2.so timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
echo "== run-2 commit again"; (cd _wt_run2 && timeout 600 python tests/tools/gpu_check.py cfg2 cfg3 2>&1 | python tests/tools/short.py)
