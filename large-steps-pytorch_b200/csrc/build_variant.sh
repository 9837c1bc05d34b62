#!/bin/bash
# ./build_variant.sh <suffix> <extra nvcc flags...>: builds ../largesteps_b200/libls_b200_<suffix>.so with the fused TUs recompiled under the flags (A/B builds)
SUF=$1; shift
NV="/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr"
mkdir -p build_$SUF
for f in ls_pcg ls_fused_a ls_fused_b ls_fused_c; do $NV "$@" -c $f.cu -o build_$SUF/$f.o 2> build_$SUF/$f.log & done; wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../largesteps_b200/libls_b200_$SUF.so build/ls_capi.o build/ls_assemble.o build/ls_order.o build/ls_spmm.o build/ls_adam.o build/ls_glue.o build_$SUF/ls_pcg.o build_$SUF/ls_fused_a.o build_$SUF/ls_fused_b.o build_$SUF/ls_fused_c.o -lcudart_static -lpthread -ldl -lrt && echo built $SUF
