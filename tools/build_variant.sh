#!/bin/bash
# Kernel experiments: build libcubit_gpu_<name>.so with extra -D flags; run with CUBIT_GPU_LIB=<path>.
# usage: tools/build_variant.sh <name> [-DFOO=1 ...]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=duckdb-cubit_b200/build/variants/$name
mkdir -p $out
for u in scan_kernel aux_kernels column_decode wah_decode cubit_gpu; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function \
    --expt-relaxed-constexpr "$@" -c duckdb-cubit_b200/csrc/$u.cu -o $out/$u.o &
done
wait
nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $out/libcubit_gpu.so $out/*.o -lpthread -ldl -lrt
echo $out/libcubit_gpu.so
