#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "snr" 2>&1 | tail -1
