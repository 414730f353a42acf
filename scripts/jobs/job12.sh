#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled" > gpurun_out/r2_job12_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job12_tests.log)
tail -3 gpurun_out/r2_job12_tests.log
for sw in "SLMM_QT_PREFETCH=0" "SLMM_QT_PREFETCH=1"; do echo "== $sw"; (env $sw timeout 600 python scripts/tile_profile.py 2>&1 | grep -v Warn | tail -9) | tee -a gpurun_out/r2_tile_profile2.log; done
