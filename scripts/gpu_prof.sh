#!/bin/bash
# ncu full capture of the render kernel for one workload (after a plain run exits 0)
W=${1:-cfg2}; TAG=${2:-prof}
python scripts/prof_run.py $W 4 > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/$TAG python scripts/prof_run.py $W 4 > gpurun_out/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/${TAG}_plain.log
