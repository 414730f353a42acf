#!/bin/bash
mkdir -p gpurun_out
export PATH=/usr/local/cuda/bin:$PATH
echo "== tiled quadform tests (TMA off)"
SLMM_TMA=0 timeout 300 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled" > gpurun_out/dbg_tiled.log 2>&1; echo "rc=$?" >> gpurun_out/dbg_tiled.log
tail -15 gpurun_out/dbg_tiled.log
echo "== tiled under memcheck (c1mini)"
SLMM_TMA=0 timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled and c1mini" > gpurun_out/dbg_tiled_memcheck.log 2>&1
grep -m1 -A25 "Invalid\|ERROR SUMMARY" gpurun_out/dbg_tiled_memcheck.log | head -60
echo "== TMA wide supernode test under memcheck"
timeout 600 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "wide_supernodes" > gpurun_out/dbg_tma_memcheck.log 2>&1
grep -m1 -B2 -A25 "Invalid\|Illegal\|ERROR SUMMARY" gpurun_out/dbg_tma_memcheck.log | head -60
