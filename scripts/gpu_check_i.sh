#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "snr" 2>&1 | tail -2
for r in 1 2 3; do echo "R7: $(timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1 | cut -c1-110)"; done
