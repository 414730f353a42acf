#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "wide_front or headline or narrow or tiled" > gpurun_out/r2_job6_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job6_tests.log)
tail -12 gpurun_out/r2_job6_tests.log
(timeout 600 python scripts/eval_breakdown.py > gpurun_out/r2_breakdown6.log 2>&1; echo "rc=$?" >> gpurun_out/r2_breakdown6.log)
grep -v Warn gpurun_out/r2_breakdown6.log | tail -32
(timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err)
tail -6 gpurun_out/r2_bench.err
head -c 7000 gpurun_out/r2_bench.json
