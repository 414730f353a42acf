#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "headline" > gpurun_out/r2_job7_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job7_tests.log)
tail -5 gpurun_out/r2_job7_tests.log
bash scripts/jobs/bench_ref.sh
