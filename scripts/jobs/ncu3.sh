#!/bin/bash
mkdir -p gpurun_out
python scripts/tile_profile.py > gpurun_out/ncu3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:quadform_tiled_kernel -s 3 -c 1 -f -o gpurun_out/r02_quadform_tiled \
    python scripts/tile_profile.py > gpurun_out/ncu3.log 2>&1
tail -4 gpurun_out/ncu3.log
ls -la gpurun_out/r02_quadform_tiled.ncu-rep
