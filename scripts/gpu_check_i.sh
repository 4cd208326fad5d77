#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_talker.py -x -q -m gpu -k "lanes or two_handles or trapped" 2>&1 | tail -2
