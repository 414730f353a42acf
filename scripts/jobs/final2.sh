#!/bin/bash
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final2_tests.log)
tail -4 gpurun_out/r2_final2_tests.log
(timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_final2_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final2_smoke.log)
tail -2 gpurun_out/r2_final2_smoke.log
(timeout 1500 python bench.py > gpurun_out/r2_final2_bench.json 2> gpurun_out/r2_final2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_final2_bench.err)
tail -3 gpurun_out/r2_final2_bench.err
head -c 600 gpurun_out/r2_final2_bench.json
