#!/bin/bash
# builds variants of libmps_b200.so with different NN scan launch shapes into profiles/tune/ (run here),
# then `python profiles/tools/nn_probe.py profiles/tune/<variant>.so` on the GPU box for each.
set -e
cd "$(dirname "$0")/../../manipulation_planning_suite_b200/csrc"
mkdir -p ../../profiles/tune
build() { # name, flags
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
     -Xcompiler -fPIC,-ffp-contract=off -cudart static -shared $2 -o ../../profiles/tune/$1.so abi.cu rollout.cu nn_scan.cu peak.cu
}
build t256_c3_b4 "-DNN_THREADS=256 -DNN_CTAS_PER_SM=3 -DNN_BODY_BATCH=4" &
build t256_c2_b4 "-DNN_THREADS=256 -DNN_CTAS_PER_SM=2 -DNN_BODY_BATCH=4" &
build t128_c6_b4 "-DNN_THREADS=128 -DNN_CTAS_PER_SM=6 -DNN_BODY_BATCH=4" &
build t256_c2_b6 "-DNN_THREADS=256 -DNN_CTAS_PER_SM=2 -DNN_BODY_BATCH=6" &
wait
build t256_c4_b3 "-DNN_THREADS=256 -DNN_CTAS_PER_SM=4 -DNN_BODY_BATCH=3" &
build t512_c2_b3 "-DNN_THREADS=512 -DNN_CTAS_PER_SM=2 -DNN_BODY_BATCH=3" &
build t256_c3_b4_cs "-DNN_THREADS=256 -DNN_CTAS_PER_SM=3 -DNN_BODY_BATCH=4 -DNN_LOAD=__ldcs" &
build t128_c8_b3 "-DNN_THREADS=128 -DNN_CTAS_PER_SM=8 -DNN_BODY_BATCH=3" &
wait
ls -la ../../profiles/tune
