#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -k "no_projection" 2>&1 | tail -3
