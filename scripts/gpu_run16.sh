set -u
timeout 300 python tests/tools/phase_diff.py cfg3 8 2>&1 | tail -4
timeout 300 python tests/tools/phase_diff.py cfg5 8 2>&1 | tail -3
timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=600 2>&1 | tail -5
