#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r2_job8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job8_tests.log)
tail -3 gpurun_out/r2_job8_tests.log
for sw in "SLMM_CHAIN_SMALL=0" "SLMM_CHAIN_SMALL=1"; do
  echo "== $sw"
  (env $sw timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|factor_at|solve_|lmul|tiled") | tee -a gpurun_out/r2_breakdown8.log
done
(timeout 900 python scripts/launch_profile.py 250000 1e-3 0.065625 > gpurun_out/r2_launch_profile8.log 2>&1; echo "rc=$?" >> gpurun_out/r2_launch_profile8.log)
grep -v Warn gpurun_out/r2_launch_profile8.log | grep -E "^==|^  [a-z_]+ +n=" 
