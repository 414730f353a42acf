#!/bin/bash
# 1-GPU bench (both arms), as the driver runs it but with fewer steps
mkdir -p gpurun_out
(timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err)
tail -12 gpurun_out/r2_bench.err
head -c 3000 gpurun_out/r2_bench.json
