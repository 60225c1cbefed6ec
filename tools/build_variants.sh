#!/bin/bash
# Developer A/B helper: builds libslide_pr variants with different CTA size / register targets
# into variants/ (git-ignored *.so); select one with SLIDE_PR_LIB=variants/libslide_pr_<tag>.so
set -e
cd "$(dirname "$0")/../slide_slam_b200/csrc"
mkdir -p ../../variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden"
for cfg in "$@"; do
  tag=$(echo "$cfg" | tr ' =-' '___' | tr -d 'D')
  $NV $cfg -c -o /tmp/spr_kernels_$tag.o spr_kernels.cu
  $NV -shared -o ../../variants/libslide_pr_$tag.so spr_host.o /tmp/spr_kernels_$tag.o spr_kernels_bound.o spr_kernels_aux.o spr_api.o -cudart static
  echo "built variants/libslide_pr_$tag.so"
done
