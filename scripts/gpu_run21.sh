set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_cfg4.py 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout=600 -k "bvh or cfg4" 2>&1 | tail -3
python scripts/prof_run.py cfg4 4 > gpurun_out/prof_cfg4_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_cfg4 python scripts/prof_run.py cfg4 4 > gpurun_out/prof_cfg4_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_cfg4_plain.log
