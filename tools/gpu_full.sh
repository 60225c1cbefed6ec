#!/bin/bash
# full GPU suite, per-step instruction counts of the three engines, bench line, reference arm
set -u
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/full_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/full_pytest.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/full_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/full_smoke.log
M=smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
for m in join exhaustive bound; do
  timeout 120 python tools/ncu_step_target.py $m > gpurun_out/full_step_$m.log 2>&1 && \
  timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/full_step_${m}_launches.csv python tools/ncu_step_target.py $m > gpurun_out/full_ncu_$m.log 2>&1
done
python tools/ncu_counts.py gpurun_out/full_step_join_launches.csv gpurun_out/full_step_exhaustive_launches.csv gpurun_out/full_step_bound_launches.csv profiles/r2_instr_counts.json > gpurun_out/full_counts.log 2>&1
cp profiles/r2_instr_counts.json gpurun_out/full_instr_counts.json
timeout 900 python bench.py > gpurun_out/full_bench.json 2> gpurun_out/full_bench.err; echo "bench rc=$?" >> gpurun_out/full_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/full_bench_ref.json 2> gpurun_out/full_bench_ref.err; echo "ref rc=$?" >> gpurun_out/full_bench_ref.err
