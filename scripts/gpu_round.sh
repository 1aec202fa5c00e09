#!/bin/bash
# Full GPU validation pass: tests, smoke, bench, ncu launch list and a full capture of the top kernel.
set -u
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi.csv
if [ "${1:-}" = "ncu" ]; then
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/plain2.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:draw_ -s 3 -c 2 -f -o gpurun_out/prof \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  python bench.py --strict --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/plain3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:draw_ -s 3 -c 1 -f -o gpurun_out/prof_strict \
      python bench.py --strict --steps 3 --warmup 3 --no-cpu-baseline --no-4k > gpurun_out/ncu_full_strict.log 2>&1
  echo "ncu full strict rc=$?"
fi
