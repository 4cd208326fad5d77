#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; tail -2 gpurun_out/r2e_bench_n2.err
python -c "
import json
j = json.loads(open('gpurun_out/r2e_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value', j['value'], 'e2e', j['e2e']['value'], 'n_gpus', j['n_gpus'])
print(json.dumps({k: v for k, v in j['config3'].items() if k != 'workload'}))
print(json.dumps({k: v for k, v in j['config4'].items() if k != 'workload'}))
print('c5', j['config5'])
"
