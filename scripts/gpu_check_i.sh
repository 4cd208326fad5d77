#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_codec.py tests/test_gpu_fullsize.py -x -q -m gpu -k "codec or snr or decode or window or chunk" 2>&1 | tail -3
timeout 900 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "snr" -s 2>&1 | grep -i "snr\|passed\|failed" | head
for r in 1 2; do
echo "simt: $(Q3TTS_CODEC_ATT_MMA=0 timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1 | cut -c1-120)"
echo "mma : $(timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1 | cut -c1-120)"
done
echo "simt T=750: $(Q3TTS_CODEC_ATT_MMA=0 timeout 300 python scripts/codec_probe.py 8 750 3 2>&1 | tail -1 | cut -c1-120)"
echo "mma  T=750: $(timeout 300 python scripts/codec_probe.py 8 750 3 2>&1 | tail -1 | cut -c1-120)"
