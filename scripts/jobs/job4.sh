#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests4.log)
tail -8 gpurun_out/r2_tests4.log
(timeout 300 python scripts/eval_breakdown.py 250000 > gpurun_out/r2_breakdown4.log 2>&1)
grep -v Warn gpurun_out/r2_breakdown4.log
(SLMM_SKINNY=0 timeout 300 python scripts/eval_breakdown.py 250000 > gpurun_out/r2_breakdown4_noskinny.log 2>&1)
grep "solve_\|lmul 16\|hess\|fixed" gpurun_out/r2_breakdown4_noskinny.log
(timeout 400 python scripts/launch_profile.py 250000 1e-3 0.065625 > gpurun_out/r2_launch_profile4.log 2>&1)
grep -v Warn gpurun_out/r2_launch_profile4.log | grep "==\|skinny\|rhs_pull\|reduce \|gemm_tma\|slowest" | cut -c1-400
