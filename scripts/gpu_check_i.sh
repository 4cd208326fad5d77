#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/codec_probe.py 64 26 2 > gpurun_out/r2_g_plain.jsonl 2> gpurun_out/r2_g.err && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_codec_final.csv python scripts/codec_probe.py 64 26 1 > gpurun_out/r2_g_ncu.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_codec_final.csv 2>/dev/null | head -6
