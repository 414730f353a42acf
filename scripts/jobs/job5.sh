#!/bin/bash
# narrow-RHS kernel tests, evaluation breakdown, then the 1-GPU bench
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_surface.py tests/test_gpu_parity.py -m gpu -x -q -k "narrow or solve or hess or reml or tiled" > gpurun_out/r2_job5_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job5_tests.log)
tail -4 gpurun_out/r2_job5_tests.log
(timeout 600 python scripts/eval_breakdown.py > gpurun_out/r2_breakdown5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_breakdown5.log)
grep -v Warn gpurun_out/r2_breakdown5.log | tail -32
(timeout 1500 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err)
tail -6 gpurun_out/r2_bench.err
head -c 6000 gpurun_out/r2_bench.json
