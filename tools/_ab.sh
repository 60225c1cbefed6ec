python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -2 gpurun_out/ab_t.log
for v in variants/*.so; do echo $v; SLIDE_PR_LIB=$v python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -1 | cut -c1-140; done
echo default; python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -1 | cut -c1-140
SLIDE_PR_TRACE=1 python tools/trace_e2e.py 2>&1 | tail -2
python tools/profile_target.py 2 > gpurun_out/pt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spr_bound -s 8 -c 1 -o gpurun_out/prof_r1_p_bound -f python tools/profile_target.py 2 > gpurun_out/ncu_p.log 2>&1
