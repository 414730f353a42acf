#!/bin/bash
# round-2 ncu evidence: launch list of one evaluation + HE, --set full of the TMA tile GEMM, the tiled quadratic-form
# pass and the narrow-RHS streaming kernels.  Every ncu command follows a plain run of the same command that exited 0.
mkdir -p gpurun_out
python scripts/gemm_lower.py 16384 16384 1024 > gpurun_out/ncu2_plain_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tiles_tma -s 1 -c 2 -f -o gpurun_out/r02_gemm_tma \
    python scripts/gemm_lower.py 16384 16384 1024 > gpurun_out/ncu2_gemm.log 2>&1
tail -3 gpurun_out/ncu2_plain_gemm.log gpurun_out/ncu2_gemm.log
python scripts/profile_step.py > gpurun_out/ncu2_plain_step.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python scripts/profile_step.py > gpurun_out/ncu2_step.log 2>&1
tail -3 gpurun_out/ncu2_plain_step.log gpurun_out/ncu2_step.log
python scripts/profile_step.py > gpurun_out/ncu2_plain_step2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:quadform_tiled_kernel|qt_gram_finish|skinny_f1|skinny_f2" -s 40 -c 12 -f -o gpurun_out/r02_quad_skinny \
    python scripts/profile_step.py > gpurun_out/ncu2_step2.log 2>&1
tail -3 gpurun_out/ncu2_step2.log
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches.csv
