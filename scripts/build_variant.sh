#!/bin/bash
# build_variant.sh NAME "-DMACRO=.. ..." : libsphmw.so with pair_ops.cu compiled with extra macros,
# stored as sph_mountain_waves_b200/build/variants/libsphmw_NAME.so (A/B runs on the GPU box)
set -e
cd "$(dirname "$0")/../sph_mountain_waves_b200"
mkdir -p build/variants
NV="nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC,-O2 -Xcudafe --diag_suppress=177 -I../include -Icsrc"
$NV $2 -Xptxas=-v -c csrc/pair_ops.cu -o build/variants/pair_ops_$1.o 2> build/variants/ptxas_$1.log
objs=""
for o in api cell_list halo slab_comm frame_async lattice frame_io grid_setup; do objs="$objs build/$o.o"; done
nvcc -ccbin /usr/bin/g++ -shared -gencode arch=compute_100a,code=sm_100a -o build/variants/libsphmw_$1.so build/variants/pair_ops_$1.o $objs -lcudart -lz -ldl
grep -A3 "k_binary_buildILi3E21B_wcsph_density_fusedLi3ELb1" build/variants/ptxas_$1.log | grep -E "Used|spill"
