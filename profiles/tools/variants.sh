#!/bin/bash
# Builds variants of libmps_b200.so that differ in compile-time switches of the rollout kernel (name=flags ...)
# into profiles/tools/_bin/ (in parallel), for profiles/tools/variant_bench.py to time on the GPU box:
#   profiles/tools/variants.sh base= noone=-DMPS32_ONE=0 ...
set -e
cd "$(dirname "$0")/../../manipulation_planning_suite_b200/csrc"
OUT=../../profiles/tools/_bin
mkdir -p $OUT
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -cudart static"
# everything but rollout.cu is shared by the variants
for f in abi nn_scan peak forest_step; do
  [ $OUT/$f.o -nt $f.cu ] || $NVCC $FLAGS -c -o $OUT/$f.o $f.cu &
done
for spec in "$@"; do
  name=${spec%%=*}; defs=${spec#*=}
  $NVCC $FLAGS $defs -c -o $OUT/rollout_$name.o rollout.cu &
done
wait
for spec in "$@"; do
  name=${spec%%=*}
  $NVCC $FLAGS -shared -o $OUT/libmps_$name.so $OUT/abi.o $OUT/nn_scan.o $OUT/peak.o $OUT/forest_step.o $OUT/rollout_$name.o
done
ls -la $OUT/*.so
