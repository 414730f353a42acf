#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_parity.py -m gpu -x -q -k "narrow or headline or lmul or wide or simulate or probe" > gpurun_out/r2_job23_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job23_tests.log)
tail -3 gpurun_out/r2_job23_tests.log
(timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|factor_at|solve_|lmul|fixed|tiled") | tee -a gpurun_out/r2_breakdown23.log
