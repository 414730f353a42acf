#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_parity.py -m gpu -x -q -k "narrow or headline or lmul or two_rank" > gpurun_out/r2_job24_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job24_tests.log)
tail -3 gpurun_out/r2_job24_tests.log
for sw in "SLMM_SKINNY_MAX=16" "SLMM_SKINNY_MAX=64"; do
  echo "== $sw"
  (env $sw timeout 600 python scripts/solve_check.py 2>&1 | grep -v Warn | tail -21) | tee -a gpurun_out/r2_check24.log
done
