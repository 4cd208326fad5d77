#!/bin/bash
set -x
mkdir -p gpurun_out
python scripts/sanitize_probe.py > gpurun_out/r2_san_plain.log 2>&1 && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_probe.py > gpurun_out/r2_san_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2_san_memcheck.log
