#!/bin/bash
# final check of the last build: codec + gemm tests, smoke, one short bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_gemm_tc.py tests/test_gpu_audio_encoder.py tests/test_gpu_speaker_encoder.py -x -q -m gpu 2>&1 | tail -2
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python bench.py --no-extras 2>/dev/null | python -c "
import json,sys
j = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', j['value'], j['e2e']['value'], j['ms_per_step'], j['codec'])"
