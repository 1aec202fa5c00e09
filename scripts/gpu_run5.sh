set -u
mkdir -p gpurun_out
timeout 300 python tests/tools/gpu_check.py cfg2 2>&1 | python tests/tools/short.py
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "benchmarked or unsplit or split_lane" --timeout=300 2>&1 | tail -5
python scripts/prof_run.py cfg2 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_pers python scripts/prof_run.py cfg2 4 > gpurun_out/prof_pers_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_plain.log
