#!/bin/bash
mkdir -p gpurun_out
(timeout 1500 python bench.py > gpurun_out/r2_final3_bench.json 2> gpurun_out/r2_final3_bench.err; echo "bench rc=$?" >> gpurun_out/r2_final3_bench.err)
grep "full fit" gpurun_out/r2_final3_bench.err | cut -c1-700
tail -2 gpurun_out/r2_final3_bench.err
head -c 300 gpurun_out/r2_final3_bench.json
