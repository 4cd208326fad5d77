#!/bin/bash
mkdir -p gpurun_out
for H in 3 4; do
timeout 1500 python bench.py --no-cpu-baseline --config4 off --config3-handles $H > gpurun_out/r2b_bench_c3_h$H.json 2> gpurun_out/r2b_bench_c3_h$H.err; tail -2 gpurun_out/r2b_bench_c3_h$H.err
python -c "
import json
j = json.loads(open('gpurun_out/r2b_bench_c3_h$H.json').read().strip().splitlines()[-1])
c = j['config3']; print('H', c['handles_per_gpu'], 'value', c['value'], 'e2e', c['e2e'], 'wall', c['wall_s'])
"
done
