python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -2 gpurun_out/ab_t.log
python tools/nopeak_timing.py 2>&1 | tail -5
SLIDE_PR_TRACE=1 python tools/trace_e2e.py 2>&1 | tail -4
