python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -5 gpurun_out/ab_t.log
python tools/quick_bench.py 2 1 2>&1 | grep cfg
