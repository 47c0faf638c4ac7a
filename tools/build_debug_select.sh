#!/bin/bash
# librsm.so with the selection kernel's globaltimer probes (rsm_debug_select); rebuild normally afterwards
set -e
cd "$(dirname "$0")/.."
C=roborts_edu_slam_b200/csrc
nvcc -t 0 -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -DRSM_SELECT_DEBUG \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math -shared -o roborts_edu_slam_b200/librsm.so $C/rsm_kernels.cu $C/rsm_score.cu $C/rsm_api.cu
