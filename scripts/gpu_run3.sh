set -u
mkdir -p gpurun_out
python scripts/prof_run.py cfg2 4 > gpurun_out/prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_new python scripts/prof_run.py cfg2 4 > gpurun_out/prof_new_ncu.log 2>&1
echo "ncu new rc=$?"; cat gpurun_out/prof_plain.log
UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_nolazy.so ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_nolazy python scripts/prof_run.py cfg2 4 > gpurun_out/prof_nolazy_ncu.log 2>&1
echo "ncu nolazy rc=$?"
