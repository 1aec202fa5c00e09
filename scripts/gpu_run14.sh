set -u
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=600 -k "lane_mapping or benchmarked or unsplit or split or mixed or interleaved" 2>&1 | tail -8
timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
timeout 300 python scripts/gpu_stride.py 2>&1 | tail -4
