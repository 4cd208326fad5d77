#!/bin/bash
# On-box: GPU tests + A/B timing of the decode-step switches.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests -q -m gpu -s --durations=8 -p no:cacheprovider > $O/r2_t4.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_t4.log
tail -4 $O/r2_t4.log
run() { name=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-extras > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err; echo "$name rc=$? $(python -c "import json;j=json.load(open('$O/r2_ab_$name.json'));print(round(j['value'],1), round(j['e2e']['value'],1), round(j['talker']['ms_per_frame_step_batch'],3), j['gpu_launches'])" 2>&1)"; }
run default Q3TTS_X=0
run dense Q3TTS_SKINNY_Q=0
run chain Q3TTS_CHAIN=1
run kv32 Q3TTS_KV_F16=0
run default2 Q3TTS_X=0
