#!/bin/bash
# Kernel experiments: builds libhc_<tag>.so variants of the product library with different -D settings of the traversal kernel.
# usage: scripts/build_variants.sh tag1 "-DHC_TRACE2_MINB=7 -DHC2_SSTK=8" tag2 "..." ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/hydracore_b200/csrc
B=/tmp/hcx/var; mkdir -p $B
FL="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 --fmad=false -std=c++17 -Xcompiler -fPIC,-fopenmp,-O3"
unset CXX CC
if [ ! -f $B/hc_path.o ] || [ $CS/hc_path.cu -nt $B/hc_path.o ] || [ $CS/hc_shade.cuh -nt $B/hc_path.o ]; then nvcc $FL -c -o $B/hc_path.o $CS/hc_path.cu & fi
if [ ! -f $B/bvh.o ] || [ $CS/bvh_builder.cpp -nt $B/bvh.o ]; then nvcc $FL -c -o $B/bvh.o $CS/bvh_builder.cpp & fi
while [ $# -gt 1 ]; do
  tag=$1; defs=$2; shift 2
  ( nvcc $FL $defs -c -o $B/api_$tag.o $CS/hc_api.cu ) &
  TAGS="$TAGS $tag"
done
wait
for tag in $TAGS; do nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $ROOT/hydracore_b200/libhc_$tag.so $B/api_$tag.o $B/hc_path.o $B/bvh.o -lgomp; done
ls -la $ROOT/hydracore_b200/libhc_*.so
