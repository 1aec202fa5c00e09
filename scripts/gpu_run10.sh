set -u
mkdir -p gpurun_out
(cd _wt_run2 && python scripts/prof_run.py cfg3 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:draw_ -s 1 -c 1 -f -o ../gpurun_out/prof_c3_run2 python scripts/prof_run.py cfg3 3 > ../gpurun_out/prof_c3_run2.log 2>&1; echo "rc=$?")
python scripts/prof_run.py cfg3 3 > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:draw_ -s 1 -c 1 -f -o gpurun_out/prof_c3_cur python scripts/prof_run.py cfg3 3 > gpurun_out/prof_c3_cur.log 2>&1; echo "rc=$?"
