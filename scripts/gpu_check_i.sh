#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python bench.py --no-cpu-baseline --config3 off --config4 off > gpurun_out/r2b_bench_c5.json 2> gpurun_out/r2b_bench_c5.err; tail -5 gpurun_out/r2b_bench_c5.err
python -c "
import json
j = json.loads(open('gpurun_out/r2b_bench_c5.json').read().strip().splitlines()[-1])
print(json.dumps(j['config5'], indent=1))
"
