#!/bin/bash
# Builds dcdf_b200/libdcdf_cuda.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC -Xptxas -v"
OUT=../libdcdf_cuda.so
mkdir -p ../_build
pids=()
for f in api_encode api_decode api_store; do
  ( $NVCC $FLAGS -c $f.cu -o ../_build/$f.o > ../_build/$f.log 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait $p || rc=1; done
if [ $rc -ne 0 ]; then cat ../_build/*.log | grep -v "^ptxas info" | head -80; exit 1; fi
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o $OUT ../_build/api_encode.o ../_build/api_decode.o ../_build/api_store.o -lcudart_static -lpthread -ldl -lrt
echo built $OUT
