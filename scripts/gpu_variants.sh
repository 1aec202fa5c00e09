#!/bin/bash
# run gpu_check for the default lib and each kernel-variant lib in uob_raytracer_b200/variants/var_*.so
echo "== default"; python tests/tools/gpu_check.py "$@" 2>&1 | python tests/tools/short.py
for f in uob_raytracer_b200/variants/var_*.so; do echo "== $f"; UOB_RT_LIB=$PWD/$f python tests/tools/gpu_check.py "$@" 2>&1 | python tests/tools/short.py; done
