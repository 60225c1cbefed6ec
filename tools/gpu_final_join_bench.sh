#!/bin/bash
# final check of the round: pair-join / cache GPU suites, launch list of one default search step (instruction counts for the
# bench line's roofline), the bench line, then an ncu --set full capture of the pair-join kernel
set -u
mkdir -p gpurun_out
timeout 260 python -m pytest tests/test_gpu_join.py tests/test_gpu_cache.py -m gpu -q -x --durations=5 -k "not config3" > gpurun_out/v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/v_pytest.log
M=smsp__inst_executed.sum,smsp__inst_executed_pipe_alu.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
timeout 100 python tools/ncu_step_target.py join > gpurun_out/v_step_join.log 2>&1 && \
timeout 200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/v_step_join_launches.csv python tools/ncu_step_target.py join > gpurun_out/v_ncu_join.log 2>&1 && \
python tools/ncu_counts.py gpurun_out/v_step_join_launches.csv profiles/r2_step_exhaustive_launches.csv profiles/r2_step_bound_launches.csv profiles/r2_instr_counts.json > gpurun_out/v_counts.log 2>&1
cp profiles/r2_instr_counts.json gpurun_out/v_instr_counts.json
timeout 600 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?" >> gpurun_out/v_bench.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:spr_join_score -c 1 -o gpurun_out/v_join_c2 python tools/ncu_step_target.py join > gpurun_out/v_ncu_full.log 2>&1
timeout 100 ncu -i gpurun_out/v_join_c2.ncu-rep --page raw --csv > gpurun_out/v_join_c2_raw.csv 2> gpurun_out/v_err1.log
