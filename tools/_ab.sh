for v in variants/*.so; do
  SLIDE_PR_LIB=$v python -m pytest tests -x -q -m gpu -k "golden or random or counts" > gpurun_out/ab_t.log 2>&1; echo "$v $(tail -1 gpurun_out/ab_t.log)"
  SLIDE_PR_LIB=$v python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -2
done
echo default; python tools/quick_bench.py 2 2>&1 | grep cfg2 | head -2
python tools/profile_target.py 2 > gpurun_out/pt.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:spr_score_lattice -s 12 -c 1 -o gpurun_out/prof_r1_k -f python tools/profile_target.py 2 > gpurun_out/ncu_k.log 2>&1
