#!/bin/bash
# run gpu_check for the default lib and each kernel-variant lib in uob_raytracer_b200/build/var_*.so
echo "== default"; python scripts/gpu_check.py "$@" 2>&1 | python scripts/short.py
for f in uob_raytracer_b200/build/var_*.so; do echo "== $f"; UOB_RT_LIB=$PWD/$f python scripts/gpu_check.py "$@" 2>&1 | python scripts/short.py; done
