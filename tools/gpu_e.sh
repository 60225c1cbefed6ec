#!/bin/bash
# bench line (new layout) + per-step instruction counts for its roofline
set -u
mkdir -p gpurun_out
M=smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 120 python tools/ncu_step_target.py exhaustive > gpurun_out/e_step_x.log 2>&1 && \
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/e_step_exhaustive_launches.csv python tools/ncu_step_target.py exhaustive > gpurun_out/e_ncu_x.log 2>&1
timeout 120 python tools/ncu_step_target.py default > gpurun_out/e_step_d.log 2>&1 && \
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/e_step_default_launches.csv python tools/ncu_step_target.py default > gpurun_out/e_ncu_d.log 2>&1
python tools/ncu_counts.py gpurun_out/e_step_exhaustive_launches.csv gpurun_out/e_step_default_launches.csv profiles/r2_instr_counts.json > gpurun_out/e_counts.log 2>&1
cp profiles/r2_instr_counts.json gpurun_out/e_instr_counts.json
timeout 600 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?" >> gpurun_out/e_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/e_bench_ref.json 2> gpurun_out/e_bench_ref.err; echo "ref rc=$?" >> gpurun_out/e_bench_ref.err
