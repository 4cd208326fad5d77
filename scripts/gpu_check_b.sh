#!/bin/bash
# On-box: GPU tests + A/B timing of the decode-step switches (chain signals, fp16 KV rings).  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -q -m gpu -x -s --durations=8 -p no:cacheprovider > $O/r2_t3.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_t3.log
tail -4 $O/r2_t3.log
run() { name=$1; shift; env "$@" python bench.py --steps 5 --warmup 3 --no-extras > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err; echo "$name rc=$? $(python -c "import json;j=json.load(open('$O/r2_ab_$name.json'));print(round(j['value'],1), round(j['e2e']['value'],1), round(j['talker']['ms_per_frame_step_batch'],3), j['gpu_launches'])" 2>&1)"; }
run default Q3TTS_X=0
run nochain Q3TTS_CHAIN=0
run kv32 Q3TTS_KV_F16=0
run nochain_kv32 Q3TTS_CHAIN=0 Q3TTS_KV_F16=0
run dense Q3TTS_SKINNY_Q=0
