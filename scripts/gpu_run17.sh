set -u
mkdir -p gpurun_out
echo "== default"; timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
echo "== 2 blocks/SM, 128 regs"; UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_m2.so timeout 300 python tests/tools/gpu_check.py head cfg2 cfg3 2>&1 | python tests/tools/short.py
python scripts/prof_run.py cfg2 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_r02_fast python scripts/prof_run.py cfg2 4 > gpurun_out/prof_r02_fast.log 2>&1
echo "ncu fast rc=$?"
python scripts/prof_run.py cfg2 4 strict > gpurun_out/prof_plain_s.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_r02_strict python scripts/prof_run.py cfg2 4 strict > gpurun_out/prof_r02_strict.log 2>&1
echo "ncu strict rc=$?"
