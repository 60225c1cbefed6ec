python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -3 gpurun_out/ab_t.log
python tools/quick_bench.py 2 2>&1 | grep cfg2
for v in variants/libslide_pr_b_SPB_CSA.so; do SLIDE_PR_LIB=$v python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -1; done
SLIDE_PR_TRACE=1 python tools/_trace.py 2>&1 | tail -4
