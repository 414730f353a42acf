#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled" > gpurun_out/r2_job13_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job13_tests.log)
tail -3 gpurun_out/r2_job13_tests.log
(timeout 600 python scripts/tile_profile.py 2>&1 | grep -v Warn | tail -9) | tee gpurun_out/r2_tile_profile3.log
