#!/bin/bash
# fused-GEMM: parity, dequantisation phase breakdown, fused-vs-dense step time
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_talker.py -x -q -m gpu > gpurun_out/r2_t7.log 2>&1; tail -3 gpurun_out/r2_t7.log
Q3TTS_SKQ_DBG=1 python scripts/skinny_trace.py 64 4 > gpurun_out/r2_trace_dbg1b.jsonl 2> gpurun_out/r2_trace_dbg1b.err
python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/r2_ab2_fused.json 2> gpurun_out/r2_ab2_fused.err
Q3TTS_SKINNY_Q=0 python bench.py --steps 6 --warmup 3 --no-extras > gpurun_out/r2_ab2_dense.json 2> gpurun_out/r2_ab2_dense.err
tail -2 gpurun_out/r2_ab2_fused.err
