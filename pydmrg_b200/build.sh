#!/bin/bash
# Build libpydmrg_b200.so in-tree for sm_100a (B200).  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")/csrc"
# -fcx-limited-range: plain complex multiply / divide in the HOST code (the m x m Hessenberg-QR of the Davidson loop runs
# 2x faster without the C99 NaN-recovery calls); device code is unaffected
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fcx-limited-range -shared \
     -I/usr/local/cuda/include ${PDMRG_NVCC_FLAGS} \
     -o ../libpydmrg_b200.so api.cu gemm_nt.cu tensor_ops.cu heff.cu krylov.cu solver.cu comm.cu davidson.cu env_cols.cu -ldl
