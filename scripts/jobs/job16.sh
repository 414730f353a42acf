#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "narrow" > gpurun_out/r2_job16_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job16_tests.log)
tail -3 gpurun_out/r2_job16_tests.log
for sw in "SLMM_FUSE_REDUCE=0" "SLMM_FUSE_REDUCE=1"; do
  echo "== $sw"
  (env $sw timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|solve_|lmul|fixed") | tee -a gpurun_out/r2_breakdown16.log
done
