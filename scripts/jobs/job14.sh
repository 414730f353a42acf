#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python scripts/he_upload_probe.py 2>&1 | grep -v Warn | tail -40) | tee gpurun_out/r2_he_upload_probe.log
(timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_job14_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_job14_tests.log)
tail -4 gpurun_out/r2_job14_tests.log
