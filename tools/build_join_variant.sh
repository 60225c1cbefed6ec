#!/bin/bash
# Developer A/B helper: builds a libslide_pr variant with other pair-join constants into variants/
# usage: build_join_variant.sh <tag> <nvcc/g++ -D flags...>     select with SLIDE_PR_LIB=variants/libslide_pr_<tag>.so
set -e
tag=$1; shift
cd "$(dirname "$0")/../slide_slam_b200/csrc"
mkdir -p ../../variants /tmp/spj_$tag
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden"
/usr/bin/g++ -O3 -mpopcnt -std=c++17 -fPIC -ffp-contract=off -fvisibility=hidden "$@" -c -o /tmp/spj_$tag/spr_host.o spr_host.cpp
$NV "$@" -c -o /tmp/spj_$tag/spr_join.o spr_join.cu
$NV "$@" -c -o /tmp/spj_$tag/spr_api.o spr_api.cu
$NV -shared -o ../../variants/libslide_pr_$tag.so /tmp/spj_$tag/spr_host.o spr_delaunay.o spr_kernels.o spr_kernels_bound.o spr_kernels_aux.o /tmp/spj_$tag/spr_join.o spr_clipper.o spr_generate.o /tmp/spj_$tag/spr_api.o -cudart static
echo "built variants/libslide_pr_$tag.so"
