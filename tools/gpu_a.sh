#!/bin/bash
# round-2 GPU call A: full GPU suite, short bench, baseline ncu captures of the kernels to optimise
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
nproc > gpurun_out/a_nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 200 python tools/profile_target.py 3 0 2 > gpurun_out/a_c3.log 2>&1
timeout 100 python tools/profile_target.py 2 0 3 exhaustive > gpurun_out/a_c2x.log 2>&1
# ncu: bound kernel (C2 default search), exact kernel (C2 exhaustive), C3 verification kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spr_bound_lattice -s 4 -c 2 -o gpurun_out/a_bound_c2 python tools/profile_target.py 2 0 2 > gpurun_out/a_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spr_score_lattice -s 12 -c 2 -o gpurun_out/a_exact_c2 python tools/profile_target.py 2 0 2 exhaustive > gpurun_out/a_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spr_score_lattice -s 4 -c 2 -o gpurun_out/a_verify_c3 python tools/profile_target.py 3 0 1 > gpurun_out/a_ncu3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_c3_launches.csv python tools/profile_target.py 3 0 1 > gpurun_out/a_ncu4.log 2>&1
ls -la gpurun_out > gpurun_out/a_ls.txt
