#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/codec_probe.py 64 26 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"codec_attention_mma_kernel|out_conv_stream_kernel" --launch-skip 7 --launch-count 2 -o gpurun_out/r2e_codec_small -f python scripts/codec_probe.py 64 26 1 > gpurun_out/r2e_ncu_small.log 2>&1; tail -2 gpurun_out/r2e_ncu_small.log
timeout 900 ncu --set full --clock-control none -k regex:"codec_attention_mma_kernel" --launch-skip 7 --launch-count 1 -o gpurun_out/r2e_codec_att750 -f python scripts/codec_probe.py 8 750 1 > gpurun_out/r2e_ncu_att750.log 2>&1; tail -1 gpurun_out/r2e_ncu_att750.log
