#!/bin/bash
# On-box: GEMM / codec / talker tests, traces of the fused kernel (TS mode), A/B timing incl. codec-only config 4.
cd "$(dirname "$0")/.."
O=gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_codec.py tests/test_gpu_talker.py tests/test_gpu_fullsize.py -q -m gpu -s --durations=5 -p no:cacheprovider > $O/r2_t6.log 2>&1; echo "pytest rc=$?" | tee -a $O/r2_t6.log
tail -4 $O/r2_t6.log
python scripts/skinny_trace.py 64 4 > $O/r2_trace_q4_ts.jsonl 2>&1; echo "trace q4 rc=$?"
run() { name=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-extras > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err; echo "$name rc=$? $(python -c "import json;j=json.load(open('$O/r2_ab_$name.json'));print(round(j['value'],1), round(j['e2e']['value'],1), round(j['talker']['ms_per_frame_step_batch'],3), round(j['codec']['samples_per_s_rank0']/1e6,1), j['gpu_launches'])" 2>&1)"; }
run default Q3TTS_X=0
run ss Q3TTS_SKQ_TS=0
run noring Q3TTS_TC_RES_RING=0
python - > $O/r2_codec4.json 2> $O/r2_codec4.err <<'PY'
import json, os, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'mlx-swift-qwen3-tts_b200')
import numpy as np
import qwen3tts_b200 as q
import bench
d = bench.ckpt_path("0.6b", 4)
out = {}
for ring in ("1", "0"):
    os.environ["Q3TTS_TC_RES_RING"] = ring
    eng = q.Engine(d, load_talker=False, codec_max_frames=3000)
    codes = np.random.default_rng(3).integers(0, 2048, size=(32, 750, 16)).astype(np.int32)
    eng.decode_chunked(codes[:4]); eng.decode_chunked(codes)
    t = eng.timing()
    out["chunked_ring" + ring] = 32 * 750 * 1920 / (t.decode_ms * 1e-3) / 1e6
    w = np.random.default_rng(4).integers(0, 2048, size=(64, 26, 16)).astype(np.int32)
    eng.decode(w); eng.decode(w)
    out["pass64x26_ms_ring" + ring] = eng.timing().decode_ms
    eng.close()
print(json.dumps(out))
PY
cat $O/r2_codec4.json; tail -2 $O/r2_codec4.err
