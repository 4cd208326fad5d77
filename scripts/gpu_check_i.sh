#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 > gpurun_out/r2e_bench_n4.json 2> gpurun_out/r2e_bench_n4.err; tail -2 gpurun_out/r2e_bench_n4.err
python -c "
import json
j = json.loads(open('gpurun_out/r2e_bench_n4.json').read().strip().splitlines()[-1])
print('N=4 value', j['value'], 'e2e', j['e2e']['value'], 'n_gpus', j['n_gpus'])
print('c3', j['config3']['value'], j['config3']['e2e'], j['config3']['result_gather']['seconds'])
print('c4', j['config4']['chunked']['samples_per_s'], j['config4']['whole']['samples_per_s'])
"
