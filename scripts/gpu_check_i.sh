#!/bin/bash
mkdir -p gpurun_out
for H in 4 8; do
Q3TTS_PDL=0 timeout 1500 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=0 headline', j['value'], j['e2e']['value'])
" 
Q3TTS_PDL=0 timeout 1500 python bench.py --no-cpu-baseline --config4 off --config5 off --config3-handles $H > gpurun_out/r2c_c3_nopdl_l$H.json 2> gpurun_out/r2c_c3_nopdl_l$H.err; tail -2 gpurun_out/r2c_c3_nopdl_l$H.err
python -c "
import json
j = json.loads(open('gpurun_out/r2c_c3_nopdl_l$H.json').read().strip().splitlines()[-1])
c = j['config3']; print('PDL=0 lanes', c['lanes_per_gpu'], 'value', c['value'], 'e2e', c['e2e'], 'inflight', j['batches_in_flight']['e2e_value'])
"
done
