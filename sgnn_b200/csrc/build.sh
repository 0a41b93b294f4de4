#!/bin/bash
# Build libsgnn_b200.so in-tree for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libsgnn_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
$NVCC $FLAGS -shared -o $OUT scan.cu grid.cu conv.cu pointwise.cu generate.cu dense.cu generator.cu conv_tc.cu conv_tc32.cu "$@" 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" || true
echo "built $OUT"
