#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -s -k "trapped" > gpurun_out/r2_trap.log 2>&1; tail -15 gpurun_out/r2_trap.log
