#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py tests/test_gpu_parity.py -m gpu -x -q -k "narrow or headline or solve or hess or symmetry or reml_fit or c1" > gpurun_out/r2_job15_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job15_tests.log)
tail -3 gpurun_out/r2_job15_tests.log
for sw in "SLMM_FUSE_REDUCE=0" "SLMM_FUSE_REDUCE=1"; do
  echo "== $sw"
  (env $sw timeout 600 python scripts/eval_breakdown.py 2>&1 | grep -v Warn | grep -E "evaluate|factor_at|solve_|lmul|tiled|fixed") | tee -a gpurun_out/r2_breakdown15.log
done
