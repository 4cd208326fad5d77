#!/bin/bash
# fused-GEMM dequantisation phase breakdown (measurement only)
set -x
mkdir -p gpurun_out
Q3TTS_SKQ_DBG=1 python scripts/skinny_trace.py 64 4 > gpurun_out/r2_trace_dbg1.jsonl 2> gpurun_out/r2_trace_dbg1.err
Q3TTS_SKQ_DBG=3 python scripts/skinny_trace.py 64 4 > gpurun_out/r2_trace_dbg3.jsonl 2>> gpurun_out/r2_trace_dbg1.err
Q3TTS_SKQ_DBG=1 python scripts/skinny_trace.py 64 8 > gpurun_out/r2_trace_dbg1_q8.jsonl 2>> gpurun_out/r2_trace_dbg1.err
tail -3 gpurun_out/r2_trace_dbg1.err
