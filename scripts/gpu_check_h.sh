#!/bin/bash
# full validation: every GPU test, smoke(), the default bench line (extras: latency, config 3, config 4)
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu > gpurun_out/r2_full_tests.log 2>&1; tail -3 gpurun_out/r2_full_tests.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 1500 python bench.py > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -2 gpurun_out/r2_bench_full.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; tail -1 gpurun_out/r2_bench_ref.err
