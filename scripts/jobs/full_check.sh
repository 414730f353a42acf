#!/bin/bash
# the whole -m gpu suite + a short bench
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2_check_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_check_tests.log)
tail -3 gpurun_out/r2_check_tests.log
(timeout 900 python bench.py --steps 5 --warmup 3 --skip-fit --skip-he --skip-cpu > gpurun_out/r2_check_bench.json 2> gpurun_out/r2_check_bench.err; echo "bench rc=$?" >> gpurun_out/r2_check_bench.err)
tail -2 gpurun_out/r2_check_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_check_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'])
"
