#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_full_tests.log 2>&1; tail -3 gpurun_out/r2b_full_tests.log
timeout 1500 python bench.py > gpurun_out/r2b_bench_full.json 2> gpurun_out/r2b_bench_full.err; tail -2 gpurun_out/r2b_bench_full.err
python -c "
import json
j = json.loads(open('gpurun_out/r2b_bench_full.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'])
print(json.dumps(j['batches_in_flight'], indent=1)); print(json.dumps({k: v for k, v in j['config3'].items() if k != 'workload'}))
print(json.dumps(j['config4']))
"
