#!/bin/bash
# On-box: fused-GEMM + audio-encoder tests, per-phase traces of both skinny kernels, A/B timing.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_audio_encoder.py tests/test_gpu_talker.py -q -m gpu -s --durations=5 -p no:cacheprovider > $O/r2_t5.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_t5.log
tail -4 $O/r2_t5.log
python scripts/skinny_trace.py 64 4 > $O/r2_trace_q4.jsonl 2>&1; echo "trace q4 rc=$?"
python scripts/skinny_trace.py 64 0 > $O/r2_trace_f16.jsonl 2>&1; echo "trace f16 rc=$?"
run() { name=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-extras > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err; echo "$name rc=$? $(python -c "import json;j=json.load(open('$O/r2_ab_$name.json'));print(round(j['value'],1), round(j['e2e']['value'],1), round(j['talker']['ms_per_frame_step_batch'],3), j['gpu_launches'])" 2>&1)"; }
run default Q3TTS_X=0
run dense Q3TTS_SKINNY_Q=0
