#!/bin/bash
# Builds libtmac_b200.so (sm_100a only) in-tree. Usage: ./build.sh [extra nvcc flags]
# Two translation units (host + gemv3/prefill kernels, gemv4 kernels) are compiled in parallel, then linked.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=${TMAC_B200_OUT:-t-mac_b200/libtmac_b200.so}
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden,-O2"
$NVCC $FLAGS -c -o "$OBJ/tmac_b200.o" t-mac_b200/csrc/tmac_b200.cu "$@" &
P1=$!
$NVCC $FLAGS -c -o "$OBJ/tmac_gemv4.o" t-mac_b200/csrc/tmac_gemv4.cu "$@" &
P2=$!
wait $P1; wait $P2
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o "$OUT" "$OBJ/tmac_b200.o" "$OBJ/tmac_gemv4.o"
echo "built $OUT"
