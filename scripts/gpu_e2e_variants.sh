#!/bin/bash
# rt_render end to end (scripts/gpu_e2e.py) for the default library and every kernel-variant library, three rounds each
for rnd in 1 2; do
  echo "== default"; python scripts/gpu_e2e.py "$@" 2>&1 | tail -1
  for f in uob_raytracer_b200/variants/var_*.so; do echo "== $f"; UOB_RT_LIB=$PWD/$f python scripts/gpu_e2e.py "$@" 2>&1 | tail -1; done
done
