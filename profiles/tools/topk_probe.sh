#!/bin/bash
# builds libmps_b200 variants with phases of nn_topk_kernel switched off (NN_TOPK_PROBE) to see where its time goes
set -e
cd "$(dirname "$0")/../../manipulation_planning_suite_b200/csrc"
OUT=../../profiles/tools/_bin
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC,-ffp-contract=off -cudart static"
for f in abi rollout peak forest_step; do nvcc $FLAGS -c -o $OUT/$f.o $f.cu & done
for v in 0 1 2 3 4 5; do nvcc $FLAGS -DNN_TOPK_PROBE=$v -c -o $OUT/nn_scan_p$v.o nn_scan.cu & done
wait
for v in 0 1 2 3 4 5; do nvcc $FLAGS -shared -o $OUT/libmps_topk_p$v.so $OUT/abi.o $OUT/rollout.o $OUT/peak.o $OUT/forest_step.o $OUT/nn_scan_p$v.o; done
