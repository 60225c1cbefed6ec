python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -2 gpurun_out/ab_t.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1_q.json 2> gpurun_out/bench_r1_q.err; tail -c 2600 gpurun_out/bench_r1_q.json; tail -3 gpurun_out/bench_r1_q.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/plain_q.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_q.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_q1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spr_bound -s 8 -c 1 -o gpurun_out/prof_r1_q_bound -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_q2.log 2>&1
python bench.py --config 3 --workload shard --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('C3 ms', d['ms_per_step'], 'best', d['best_num_inliers'], 'launches', d['gpu_launches'])"
python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-160
python bench.py --config 5 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-260
