#!/bin/bash
# Build a diagnostic variant of libd3il.so into profiles/build/<name>/libd3il.so (git-ignored; travels with gpurun).
#   profiles/build_variant.sh timing0 '-DD3IL_PHASE_TIMING -DD3IL_PHASE_BLOCK=0'
# Scripts pick it up with D3IL_VARIANT=<name> (profiles/_variant.py sets lib.SO_PATH before the first load).
set -e
NAME=$1; EXTRA=$2; MAXREG=${3:-120}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/profiles/build/$NAME
mkdir -p "$OUT"
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -ftz=true -prec-div=false -prec-sqrt=false -Xcompiler -fPIC -Xptxas -v $EXTRA"
cd "$ROOT/d3il_b200/csrc"
$NVCC $FLAGS -maxrregcount=$MAXREG -c -o "$OUT/env.o" d3il_kernels_env.cu 2> "$OUT/ptxas_env.log" &
$NVCC $FLAGS -c -o "$OUT/capi.o" d3il_capi.cu 2> "$OUT/ptxas_capi.log" &
wait
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libd3il.so" "$OUT/env.o" "$OUT/capi.o"
rm -f "$OUT"/*.o
grep -hE "error|registers" "$OUT"/ptxas_*.log | sort | uniq -c | sort -rn | head -6
