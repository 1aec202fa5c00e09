set -u
for v in "" f32x2 hcfast both ""; do
  echo "== current $v"; L=""; [ -n "$v" ] && L=$PWD/uob_raytracer_b200/variants/var_$v.so
  env ${L:+UOB_RT_LIB=$L} timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
done
timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
