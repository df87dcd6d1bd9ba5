#!/bin/bash
# Build the C-ABI CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU).
set -e
cd "$(dirname "$0")/bodyfitting_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC \
     -o ../libbodyfit_b200.so bf_api.cu -ldl "$@"
