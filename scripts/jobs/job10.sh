#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled or upload_columns or reml or c1" > gpurun_out/r2_job10_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job10_tests.log)
tail -3 gpurun_out/r2_job10_tests.log
(timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|factor_at|solve_|lmul|tiled|hess|quadform") | tee gpurun_out/r2_breakdown10.log
