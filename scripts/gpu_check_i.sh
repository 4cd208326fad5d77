#!/bin/bash
# call B: ncu --set full of the decode-step GEMM at 64 rows (eager frame steps so every launch is visible)
set -x
mkdir -p gpurun_out
python scripts/frame_profile.py 64 2 > gpurun_out/r2_fp_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:tc_skinny_kernel --launch-skip 300 -c 10 -o gpurun_out/r2_skinny64 -f python scripts/frame_profile.py 64 2 > gpurun_out/r2_ncu_skinny.log 2>&1
tail -2 gpurun_out/r2_ncu_skinny.log; ls -la gpurun_out/r2_skinny64.ncu-rep
