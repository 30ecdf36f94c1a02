#!/bin/bash
# Diagnostics: variants of the MPPI translation unit (register bound of the two-rollouts-per-thread K1 instantiations x form of the
# segment-sum accumulation) linked against the other objects of the regular build -> control_toolkit_b200/variants/libctk_<name>.so,
# selected at run time with CTK_LIB=<path> (bench A/B on one box).   bash tools/build_variants.sh
set -eu
cd "$(dirname "$0")/.."
python -m control_toolkit_b200.build > /dev/null
B=control_toolkit_b200/build; V=control_toolkit_b200/variants; mkdir -p $V
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I control_toolkit_b200/csrc"
build_one() {  # name, extra flags
  nvcc $FLAGS $2 -c control_toolkit_b200/csrc/ctk_mppi.cu -o $V/mppi_$1.o
  nvcc -shared -o $V/libctk_$1.so $V/mppi_$1.o $B/ctk_engine.o $B/ctk_cem.o $B/ctk_rpgd.o $B/ctk_batch.o $B/ctk_gru.o $B/ctk_env.o -gencode arch=compute_100a,code=sm_100a -lcudart
  rm -f $V/mppi_$1.o
  echo built $1
}
build_one kahan1024 "-DCTK_K1_MAXT2=1024" &
build_one plain896 "-DCTK_K1_PLAIN_SUM" &
build_one dsum896 "-DCTK_K1_DSUM" &
wait
ls -la $V
