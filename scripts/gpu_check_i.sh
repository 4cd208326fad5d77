#!/bin/bash
# final state of the round: every GPU test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2f_full_tests.log 2>&1; tail -3 gpurun_out/r2f_full_tests.log
