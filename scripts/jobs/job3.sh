#!/bin/bash
# GPU job: full -m gpu suite, stage breakdown, launch profile at the 250K config
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log)
(timeout 300 python scripts/eval_breakdown.py 250000 > gpurun_out/r2_breakdown3.log 2>&1)
(timeout 400 python scripts/launch_profile.py 250000 1e-3 0.065625 > gpurun_out/r2_launch_profile3.log 2>&1)
tail -5 gpurun_out/r2_tests3.log
grep -v Warn gpurun_out/r2_breakdown3.log
