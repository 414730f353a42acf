#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_parity.py -m gpu -x -q -k "narrow or headline or hess or c1_full or reml or lmul" > gpurun_out/r2_job22_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job22_tests.log)
tail -3 gpurun_out/r2_job22_tests.log
for sw in "SLMM_F1_DMMA=0" "SLMM_F1_DMMA=1"; do
  echo "== $sw"
  (env $sw timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|solve_|lmul 16|fixed") | tee -a gpurun_out/r2_breakdown22.log
done
