#!/bin/bash
# fused residual unit: parity, then codec pass timing; fused GEMM fold fix A/B
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py -x -q -m gpu -k "tensor_core or fused" > gpurun_out/r2_t8.log 2>&1; tail -5 gpurun_out/r2_t8.log
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "codec" -s > gpurun_out/r2_t8b.log 2>&1; tail -5 gpurun_out/r2_t8b.log
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/r2_f_unit.json 2> gpurun_out/r2_f_unit.err
Q3TTS_CODEC_UNIT=0 timeout 600 python bench.py --steps 4 --warmup 3 --no-extras > gpurun_out/r2_f_nounit.json 2> gpurun_out/r2_f_nounit.err
timeout 600 python bench.py --steps 4 --warmup 3 --no-extras --packed-gemm 1 > gpurun_out/r2_f_packed.json 2> gpurun_out/r2_f_packed.err
timeout 900 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -k "packed" > gpurun_out/r2_t8c.log 2>&1; tail -3 gpurun_out/r2_t8c.log
