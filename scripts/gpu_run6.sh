set -u
mkdir -p gpurun_out
echo "== persistent"; timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
echo "== one block per tile"; UOB_RT_GRID=tiles timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
timeout 300 python scripts/gpu_stride.py 2>&1 | tail -5
timeout 300 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "benchmarked or unsplit or split_lane" --timeout=300 2>&1 | tail -5
python scripts/prof_run.py cfg2 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_pers2 python scripts/prof_run.py cfg2 4 > gpurun_out/prof_pers2_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_plain.log
