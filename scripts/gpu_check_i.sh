#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_fullsize.py -x -q -m gpu -k "tensor_core or fused or codec" > gpurun_out/r2_t14.log 2>&1; tail -2 gpurun_out/r2_t14.log
Q3TTS_CODEC_UNIT_TRACE=gpurun_out/r2_unit_trace96_tps2.json python scripts/codec_probe.py 64 26 3 | tail -n 1
python scripts/unit_trace.py gpurun_out/r2_unit_trace96_tps2.json | tail -3
Q3TTS_CODEC_UNIT_TPS=1 python scripts/codec_probe.py 64 26 3 | tail -n 1
