#!/bin/bash
# tools/build_variant.sh <name> [nvcc flags...]  ->  yocto-hair_b200/csrc/variants/libv_<name>.so
set -e
cd "$(dirname "$0")/../yocto-hair_b200/csrc"
mkdir -p variants
name=$1; shift
for f in $(ls yh_*.cu | sed "s/\.cu$//"); do
  nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo --fmad=false -Xcompiler -fPIC,-ffp-contract=off -Xptxas -v "$@" -c $f.cu -o variants/$name.$f.o 2> variants/$name.$f.log &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libv_$name.so variants/$name.yh_*.o -lcudart
grep -A1 "k_intersectENS" variants/$name.yh_trace.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '; echo " <- $name"
