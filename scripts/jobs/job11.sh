#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python scripts/tile_profile.py 2>&1 | grep -v Warn | tail -12) | tee gpurun_out/r2_tile_profile.log
