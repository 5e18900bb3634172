#!/bin/bash
# usage: tools/mkvariant.sh <name> [-DFLAG ...]   -- builds variants/lib_<name>.so with extra nvcc flags (tuning aid)
set -e
N=$1; shift
cd "$(dirname "$0")/../feddlib_b200/csrc"
mkdir -p ../../variants /tmp/fbv_$N
for f in core tables api csrops; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f.cu -o /tmp/fbv_$N/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/lib_$N.so /tmp/fbv_$N/core.o /tmp/fbv_$N/tables.o /tmp/fbv_$N/api.o /tmp/fbv_$N/csrops.o _build/halo.o -lcudart
echo built variants/lib_$N.so
