#!/bin/bash
mkdir -p gpurun_out
for sw in "" "SLMM_SKINNY=0" "SLMM_TMA=0"; do
  (env $sw timeout 600 python scripts/solve_check.py 2>&1 | grep -v Warn | tail -14) | tee -a gpurun_out/r2_check1.log
done
