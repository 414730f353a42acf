#!/bin/bash
# full GPU suite + smoke + the 1-GPU bench on the final state
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_tests.log)
tail -4 gpurun_out/r2_final_tests.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final_smoke.log)
tail -3 gpurun_out/r2_final_smoke.log
(timeout 1500 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$?" >> gpurun_out/r2_final_bench.err)
tail -5 gpurun_out/r2_final_bench.err
head -c 1500 gpurun_out/r2_final_bench.json
