python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -2 gpurun_out/ab_t.log
python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -2
SLIDE_PR_TRACE=1 python tools/_trace.py 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench value %.3e ms %.3f e2e %.3e ms %.3f host %.3f'%(d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['host_index_build_ms']))"
