#!/bin/bash
# Build kernel-variant libraries uob_raytracer_b200/variants/var_<name>.so for A/B runs (scripts/gpu_variants.sh):
#   scripts/build_variants.sh name1="-DFLAG ..." name2="..."
set -e
for spec in "$@"; do
  name="${spec%%=*}"; defs="${spec#*=}"
  echo "== var_$name: $defs"
  mkdir -p uob_raytracer_b200/variants; UOB_BUILD_DIR=$PWD/uob_raytracer_b200/build/obj_$name UOB_RT_LIB=$PWD/uob_raytracer_b200/variants/var_$name.so UOB_NVCC_DEFS="$defs" \
    python -m uob_raytracer_b200.build --force > /dev/null; rm -rf $PWD/uob_raytracer_b200/build/obj_$name
done
ls -la uob_raytracer_b200/variants/var_*.so
