#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/codec_probe.py 64 26 1 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:codec_unit_kernel --launch-skip 2 --launch-count 2 -o gpurun_out/r2c_codec_unit -f python scripts/codec_probe.py 64 26 1 > gpurun_out/r2c_ncu_unit.log 2>&1; tail -2 gpurun_out/r2c_ncu_unit.log
