python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -3 gpurun_out/ab_t.log
python tools/quick_bench.py 2 1 2>&1 | grep cfg
python tools/profile_target.py 2 > gpurun_out/pt.log 2>&1; tail -3 gpurun_out/pt.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_m.csv python tools/profile_target.py 2 > gpurun_out/ncu_m.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spr_bound -s 4 -c 1 -o gpurun_out/prof_r1_bound -f python tools/profile_target.py 2 > gpurun_out/ncu_b.log 2>&1
