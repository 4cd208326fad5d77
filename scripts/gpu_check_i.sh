#!/bin/bash
mkdir -p gpurun_out
for S in 128 256 512; do echo "strip=$S"; Q3TTS_OUT_STRIP=$S timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1; done
echo "strip=256"; Q3TTS_OUT_STRIP=256 timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1
echo "strip=128"; Q3TTS_OUT_STRIP=128 timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1
