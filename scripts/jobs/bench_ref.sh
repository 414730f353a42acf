#!/bin/bash
mkdir -p gpurun_out
(timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo "rc=$?" >> gpurun_out/r2_bench_ref.err)
tail -8 gpurun_out/r2_bench_ref.err
head -c 2500 gpurun_out/r2_bench_ref.json
