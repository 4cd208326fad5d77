#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_speaker_encoder.py -x -q -m gpu -s > gpurun_out/r2_spk.log 2>&1; tail -15 gpurun_out/r2_spk.log
