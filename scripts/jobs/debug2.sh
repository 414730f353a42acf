#!/bin/bash
mkdir -p gpurun_out
echo "== strided / multi-op GEMM cases, one process each"
for i in 0 1 2 3 4 5 6; do
  timeout 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "strided_multi_op" 2>&1 > gpurun_out/dbg_case_all.log
  break
done
python - <<'PY'
import subprocess, sys
ids = subprocess.run([sys.executable, "-m", "pytest", "tests/test_gpu_parity.py", "-m", "gpu", "-q", "--collect-only", "-k", "strided_multi_op"], capture_output=True, text=True).stdout.split("\n")
ids = [l.strip() for l in ids if "strided_multi_op[" in l]
for tid in ids:
    r = subprocess.run([sys.executable, "-m", "pytest", tid, "-m", "gpu", "-q", "-x"], capture_output=True, text=True, timeout=180)
    tail = [l for l in r.stdout.split("\n") if "passed" in l or "failed" in l or "Error" in l or "error" in l][-3:]
    print(tid.split("::")[-1], "rc", r.returncode, tail, flush=True)
PY
echo "== tiled quadform test with device asserts"
SLMM_TMA=0 timeout 300 python -m pytest tests/test_gpu_surface.py -m gpu -x -q -k "tiled" > gpurun_out/dbg_tiled2.log 2>&1
grep -i "assert\|Assertion" gpurun_out/dbg_tiled2.log | sort | uniq -c | head -20
tail -5 gpurun_out/dbg_tiled2.log
