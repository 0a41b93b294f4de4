#!/bin/bash
# Build libsgnn_b200.so in-tree for sm_100a (cross-compiles without a GPU).  One nvcc per source file, in parallel.
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libsgnn_b200.so
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v"
SRCS="scan.cu grid.cu conv.cu pointwise.cu generate.cu dense.cu generator.cu conv_tc.cu conv_tc32.cu conv_ur.cu conv_urc.cu conv_sp.cu mcubes.cu"
OBJ=$(mktemp -d)
trap 'rm -rf "$OBJ"' EXIT
: > build.log
fail=0
for s in $SRCS; do
  ( $NVCC $FLAGS "$@" -c "$s" -o "$OBJ/${s%.cu}.o" > "$OBJ/${s%.cu}.log" 2>&1 || touch "$OBJ/${s%.cu}.failed" ) &
done
wait
for s in $SRCS; do
  cat "$OBJ/${s%.cu}.log" >> build.log
  if [ -e "$OBJ/${s%.cu}.failed" ]; then fail=1; fi
done
if [ "$fail" = "1" ]; then cat build.log; exit 1; fi
OBJS=""
for s in $SRCS; do OBJS="$OBJS $OBJ/${s%.cu}.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT $OBJS >> build.log 2>&1 || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" || true
echo "built $OUT"
