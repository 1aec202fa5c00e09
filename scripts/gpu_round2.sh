#!/bin/bash
# Round-2 GPU validation pass: all GPU tests (no -x: every failure is wanted), smoke, both bench arms, optional ncu.
set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -rA --timeout=900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" gpurun_out/pytest_gpu.log | tail -30
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,temperature.gpu --format=csv > gpurun_out/smi.csv
if [ "${1:-}" = "ncu" ]; then
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-side > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  python scripts/prof_run.py cfg2 4 > gpurun_out/prof_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof python scripts/prof_run.py cfg2 4 > gpurun_out/prof_ncu.log 2>&1
  echo "ncu full rc=$?"
  python scripts/prof_run.py cfg2 4 strict > gpurun_out/prof_strict_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_strict python scripts/prof_run.py cfg2 4 strict > gpurun_out/prof_strict_ncu.log 2>&1
  echo "ncu full strict rc=$?"
  python scripts/prof_run.py cfg4 4 > gpurun_out/prof_cfg4_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:draw_ -s 2 -c 1 -f -o gpurun_out/prof_cfg4 python scripts/prof_run.py cfg4 4 > gpurun_out/prof_cfg4_ncu.log 2>&1
  echo "ncu full cfg4 rc=$?"
fi
