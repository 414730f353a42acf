#!/bin/bash
# --set full of the tensor-pipe streaming kernels of the narrow solves (the 12-column fixed-effect solve of one evaluation)
mkdir -p gpurun_out
python scripts/profile_step.py > gpurun_out/ncu4_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:skinny_f1_dmma|skinny_f2_dmma" -s 60 -c 16 -f -o gpurun_out/r02_skinny_dmma \
    python scripts/profile_step.py > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
ls -la gpurun_out/r02_skinny_dmma.ncu-rep
