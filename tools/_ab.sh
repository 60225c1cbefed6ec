ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_c3.csv python tools/profile_target.py 3 20000 1 > gpurun_out/ncu_c3.log 2>&1
tail -3 gpurun_out/ncu_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c5.csv python bench.py --config 5 --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_c5.log 2>&1
tail -2 gpurun_out/ncu_c5.log | cut -c1-300
