#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -k "lanes or two_handles" 2>&1 | tail -5
timeout 1500 python bench.py --no-cpu-baseline --config4 off --config5 off > gpurun_out/r2b_bench_lanes.json 2> gpurun_out/r2b_bench_lanes.err; tail -3 gpurun_out/r2b_bench_lanes.err
python -c "
import json
j = json.loads(open('gpurun_out/r2b_bench_lanes.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'])
print(json.dumps({k: v for k, v in j['config3'].items() if k not in ('workload',)}))
"
