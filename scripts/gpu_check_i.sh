#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/two_chain_probe.py 256 2,4 > gpurun_out/r2c_twochain256.log 2>&1; tail -3 gpurun_out/r2c_twochain256.log
timeout 900 python scripts/two_chain_probe.py 512 4,8 > gpurun_out/r2c_twochain512.log 2>&1; tail -3 gpurun_out/r2c_twochain512.log
