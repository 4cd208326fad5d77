#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -k "two_handles or equals_singles or trapped" 2>&1 | tail -3
timeout 1500 python bench.py --no-cpu-baseline --config4 off --config5 off > gpurun_out/r2b_bench_clone.json 2> gpurun_out/r2b_bench_clone.err; tail -3 gpurun_out/r2b_bench_clone.err
python -c "
import json
j = json.loads(open('gpurun_out/r2b_bench_clone.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'])
print(json.dumps({k: v for k, v in j['batches_in_flight'].items() if k != 'what'})); print(json.dumps({k: v for k, v in j['config3'].items() if k not in ('workload',)}))
"
