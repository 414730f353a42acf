#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_job9_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job9_tests.log)
tail -4 gpurun_out/r2_job9_tests.log
(timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|factor_at|solve_|lmul|tiled|hess") | tee gpurun_out/r2_breakdown9.log
