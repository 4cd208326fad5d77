#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/skinny_trace.py 64 0 > gpurun_out/r2_trace_f16_v4.jsonl 2> gpurun_out/r2_g.err
python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/r2_k_bench.json 2> gpurun_out/r2_k_bench.err
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_talker.py -x -q -m gpu > gpurun_out/r2_t15.log 2>&1; tail -2 gpurun_out/r2_t15.log
