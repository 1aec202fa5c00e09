set -u
echo "== run-2 commit (673335b)"; (cd _wt_run2 && timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py; timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4)
echo "== current"; timeout 600 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py; timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x --timeout=300 -k "not fuzz" 2>&1 | tail -4
FUZZ_DUMP=a timeout 300 python tests/tools/fuzz_diag.py 108 227 66 81
