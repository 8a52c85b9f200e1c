#!/bin/bash
# tools/build_variant.sh NAME "-DPB2_X=.. ..."  ->  build/libpbrt_b200_NAME.so (tuning builds, selected with PB2_LIB=...)
set -e
cd "$(dirname "$0")/../pbrt-rs_b200"
NAME=$1; DEFS=$2
OUT=../build/var_$NAME; mkdir -p $OUT
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-std=c++17 -O3 $ARCH -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-pthread $DEFS"
for f in api api_path kernels_traverse wavefront wavefront_volpath light_distrib bvh_hlbvh; do $NVCC $FLAGS -c csrc/$f.cu -o $OUT/$f.o & done
for t in 0 1; do for g in 0 1; do $NVCC $FLAGS -DPB2_SHADE_TABLES=$t -DPB2_SHADE_SG=$g -c csrc/wavefront_shade.cu -o $OUT/wavefront_shade_$t$g.o & done; done
for f in bvh_build camera_host; do $NVCC $FLAGS -x cu -c csrc/$f.cpp -o $OUT/$f.o & done
wait
ld -r -b binary -z noexecstack -o $OUT/sobol_tables.o data/sobol_tables.bin
$NVCC -shared $ARCH -o ../build/libpbrt_b200_$NAME.so $OUT/*.o -ldl
echo built build/libpbrt_b200_$NAME.so
