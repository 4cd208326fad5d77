#!/bin/bash
set -x
mkdir -p gpurun_out
for d in 4; do
Q3TTS_CODEC_UNIT_DBG=$d Q3TTS_CODEC_UNIT_TRACE=gpurun_out/r2_unit_trace96_dbg$d.json python scripts/codec_probe.py 64 26 2 > gpurun_out/r2_g_dbg$d.jsonl 2> gpurun_out/r2_g.err
python scripts/unit_trace.py gpurun_out/r2_unit_trace96_dbg$d.json | tail -7
tail -n 1 gpurun_out/r2_g_dbg$d.jsonl
done
