#!/bin/bash
set -x
mkdir -p gpurun_out
Q3TTS_CODEC_UNIT_TRACE=gpurun_out/r2_unit_trace96_nodiv.json python scripts/codec_probe.py 64 26 3 > gpurun_out/r2_g_nodiv.jsonl 2> gpurun_out/r2_g.err
python scripts/unit_trace.py gpurun_out/r2_unit_trace96_nodiv.json | tail -4
tail -n 1 gpurun_out/r2_g_nodiv.jsonl
Q3TTS_CODEC_UNIT=0 python scripts/codec_probe.py 64 26 3 | tail -n 1
timeout 1500 python -m pytest tests/test_gpu_codec.py tests/test_gpu_gemm_tc.py tests/test_gpu_talker.py -x -q -m gpu > gpurun_out/r2_t11.log 2>&1; tail -3 gpurun_out/r2_t11.log
python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/r2_h_bench.json 2> gpurun_out/r2_h_bench.err; tail -2 gpurun_out/r2_h_bench.err
python scripts/skinny_trace.py 64 0 > gpurun_out/r2_trace_f16_nodiv.jsonl 2>> gpurun_out/r2_g.err
