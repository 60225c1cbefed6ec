#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/s_ab.log
for lib in slide_slam_b200/libslide_pr.so variants/libslide_pr_carve84.so variants/libslide_pr_carve100.so; do
SLIDE_PR_LIB=$lib timeout 200 python tools/ab_search.py 2 10 4 >> gpurun_out/s_ab.log 2>&1
SLIDE_PR_LIB=$lib timeout 300 python tools/ab_search.py 3 2 4 >> gpurun_out/s_ab.log 2>&1
done
