#!/bin/bash
# A/B of pair-join kernel build variants (variants/libslide_pr_<tag>.so, tools/build_join_variant.sh) on configs 2 / 3 / 5,
# then the fast part of the pair-join GPU suite against the variant named by $1
set -u
mkdir -p gpurun_out
: > gpurun_out/u_ab.log
for tag in c1 c2 c3 c4 c2q c3q c4l; do
  SLIDE_PR_LIB=variants/libslide_pr_$tag.so timeout 120 python tools/ab_search.py 2 20 4 >> gpurun_out/u_ab.log 2>&1
done
for tag in c2 c3 c4 c2q c3q c4l; do
  SLIDE_PR_LIB=variants/libslide_pr_$tag.so timeout 120 python tools/ab_search.py 3 3 4 >> gpurun_out/u_ab.log 2>&1
done
for tag in c2 c4 c3q; do
  echo "== $tag" >> gpurun_out/u_ab.log; SLIDE_PR_LIB=variants/libslide_pr_$tag.so timeout 120 python tools/ncu_cfg_target.py 5 >> gpurun_out/u_ab.log 2>&1
done
SLIDE_PR_LIB=variants/libslide_pr_${1:-c2}.so timeout 240 python -m pytest tests/test_gpu_join.py -m gpu -q -x --durations=5 \
  -k "golden or indoor or slices or random or dense or edge or other" > gpurun_out/u_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/u_pytest.log
