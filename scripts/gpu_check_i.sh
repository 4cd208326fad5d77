#!/bin/bash
# full validation: every GPU test, smoke(), the default bench line (all extras), the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2c_full_tests.log 2>&1; tail -3 gpurun_out/r2c_full_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2c_smoke.log 2>&1; tail -1 gpurun_out/r2c_smoke.log
timeout 1500 python bench.py > gpurun_out/r2c_bench_full.json 2> gpurun_out/r2c_bench_full.err; tail -2 gpurun_out/r2c_bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_ref.json 2> gpurun_out/r2c_bench_ref.err
python -c "
import json
j = json.loads(open('gpurun_out/r2c_bench_full.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'launches', j['gpu_launches'])
print('inflight', j['batches_in_flight']['e2e_value']); print('c3', j['config3']['value'], j['config3']['e2e'])
print('c4', j['config4']['chunked']['samples_per_s'], j['config4']['whole']['samples_per_s'])
print('c5', j['config5']['custom_voice']['value'], j['config5']['icl_clone']['value'], j['config5']['encode_reference_audio'], j['config5']['extract_speaker_embedding'])
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
r = json.loads(open('gpurun_out/r2c_bench_ref.json').read().strip().splitlines()[-1]); print('ref', r['value'])
"
