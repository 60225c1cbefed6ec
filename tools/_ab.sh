python -m pytest tests -x -q -m gpu > gpurun_out/ab_t.log 2>&1; tail -3 gpurun_out/ab_t.log
python tools/_nopeak.py 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
