#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_triangle_pipeline.py tests/test_gpu_clipper.py tests/test_gpu_fullsize.py::test_config5_streaming_queries_against_50000_landmarks_full_size "tests/test_gpu_parity.py::test_triangle_matching_matches_oracle" -m gpu -q --durations=10 > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
