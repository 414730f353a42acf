#!/bin/bash
# the 1-GPU bench, default arguments (as the driver runs it)
mkdir -p gpurun_out
(timeout 1500 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?" >> gpurun_out/r2_bench.err)
grep "full fit" gpurun_out/r2_bench.err | cut -c1-400
tail -2 gpurun_out/r2_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r2_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'fit', d['full_fit']['wall_s'], d['full_fit']['nesdis_fast']['wall_s'], 'he', d['he']['device_ms'], d['he']['e2e_s'])
"
