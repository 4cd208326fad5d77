#!/bin/bash
# full validation: every GPU test, smoke(), the default bench line (all extras), the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2e_full_tests.log 2>&1; tail -3 gpurun_out/r2e_full_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2e_smoke.log 2>&1; tail -1 gpurun_out/r2e_smoke.log
timeout 1500 python bench.py > gpurun_out/r2e_bench_full.json 2> gpurun_out/r2e_bench_full.err; tail -2 gpurun_out/r2e_bench_full.err
python -c "
import json
j = json.loads(open('gpurun_out/r2e_bench_full.json').read().strip().splitlines()[-1])
print('value', j['value'], 'e2e', j['e2e']['value'], 'roof', j['roofline']['frac'], 'launches', j['gpu_launches'], 'ms/step', j['ms_per_step'])
print('codec', j['codec']); print('talker', j['talker']['ms_per_frame_step_batch'], j['talker']['prefill_ms_per_step'])
print('inflight', j['batches_in_flight']['e2e_value']); print('c3', j['config3']['value'], j['config3']['e2e'])
print('c4', j['config4']['chunked']['samples_per_s'], j['config4']['whole']['samples_per_s'], j['config4']['chunked']['e2e_samples_per_s'], j['config4']['whole']['e2e_samples_per_s'], j['config4']['chunked']['tensor_frac_of_peak'], j['config4']['whole']['tensor_frac_of_peak'])
print('c5', j['config5']['custom_voice']['value'], j['config5']['icl_clone']['value'], j['config5']['encode_reference_audio']['device_ms'], j['config5']['extract_speaker_embedding']['device_ms'])
print('lat', j['latency']['batch1_ms_per_frame'], j['latency']['time_to_first_chunk_ms'])
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'], 'clocks', j['clocks'])
"
