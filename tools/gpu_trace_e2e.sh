#!/bin/bash
# per-phase host timings of findTransformation on two alternating config-2 pairs
mkdir -p gpurun_out
SLIDE_PR_TRACE=1 timeout 120 python tools/trace_e2e.py > gpurun_out/w_trace.log 2>&1; echo "rc=$?" >> gpurun_out/w_trace.log
