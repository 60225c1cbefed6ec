python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -3 gpurun_out/ab_t.log
SLIDE_PR_TRACE=1 python tools/_trace.py 2>&1 | tail -6
