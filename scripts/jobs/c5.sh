#!/bin/bash
mkdir -p gpurun_out
(timeout 3000 python scripts/run_c5.py > gpurun_out/r2_c5.log 2>&1; echo "rc=$?" >> gpurun_out/r2_c5.log)
grep -v Warn gpurun_out/r2_c5.log | tail -12 | cut -c1-1500
