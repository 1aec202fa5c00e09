set -u
echo "== run-2 commit (673335b)"; (cd _wt_run2 && timeout 600 python tests/tools/gpu_check.py head 2>&1 | python tests/tools/short.py; timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4)
for v in "" nobd coop4; do
  echo "== current $v"; L=""; [ -n "$v" ] && L=$PWD/uob_raytracer_b200/variants/var_$v.so
  env ${L:+UOB_RT_LIB=$L} timeout 600 python tests/tools/gpu_check.py head cfg2 2>&1 | python tests/tools/short.py; env ${L:+UOB_RT_LIB=$L} timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=300 -k "split or mixed or gate or interleaved or fuzz" 2>&1 | tail -4
