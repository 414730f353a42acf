#!/bin/bash
# GPU job: TMA GEMM self-test first (falls back to SLMM_TMA=0 for the rest if it fails or hangs), then the full
# -m gpu suite, stage breakdown and launch profile at the 250K config
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k dmma_gemm > gpurun_out/r2_tma_selftest.log 2>&1
rc=$?
echo "selftest rc=$rc" >> gpurun_out/r2_tma_selftest.log
tail -3 gpurun_out/r2_tma_selftest.log
if [ $rc -ne 0 ]; then export SLMM_TMA=0; echo "TMA self-test failed: SLMM_TMA=0 for the rest"; fi
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests3.log)
(timeout 300 python scripts/eval_breakdown.py 250000 > gpurun_out/r2_breakdown3.log 2>&1)
(SLMM_TMA=0 timeout 300 python scripts/eval_breakdown.py 250000 > gpurun_out/r2_breakdown3_notma.log 2>&1)
(timeout 400 python scripts/launch_profile.py 250000 1e-3 0.065625 > gpurun_out/r2_launch_profile3.log 2>&1)
tail -5 gpurun_out/r2_tests3.log
grep -v Warn gpurun_out/r2_breakdown3.log
grep -v Warn gpurun_out/r2_breakdown3_notma.log | head -4
