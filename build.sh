#!/bin/bash
# Builds libtmac_b200.so (sm_100a only) in-tree. Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC,-fvisibility=hidden,-O2 -shared -cudart static \
  -o t-mac_b200/libtmac_b200.so t-mac_b200/csrc/tmac_b200.cu "$@"
echo "built t-mac_b200/libtmac_b200.so"
