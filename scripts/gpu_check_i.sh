#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2c_short_plain.json 2> gpurun_out/r2c_short_plain.err && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2c_launches_bench_short.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/r2c_short_ncu.out 2>&1
tail -1 gpurun_out/r2c_short_plain.json | cut -c1-200
timeout 300 python scripts/codec_probe.py 64 26 1 > /dev/null 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2c_launches_codec.csv python scripts/codec_probe.py 64 26 1 > gpurun_out/r2c_codec_ncu.out 2>&1
wc -l gpurun_out/r2c_launches_bench_short.csv gpurun_out/r2c_launches_codec.csv
