#!/bin/bash
mkdir -p gpurun_out
B=$PWD/mlx-swift-qwen3-tts_b200/qwen3tts_b200/libq3_base.so
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_fullsize.py -x -q -m gpu -k "tc_ or snr" 2>&1 | tail -2
for r in 1 2; do
echo "base: $(Q3TTS_LIB=$B timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1 | cut -c1-120)"
echo "new : $(timeout 300 python scripts/codec_probe.py 64 26 3 2>&1 | tail -1 | cut -c1-120)"
done
