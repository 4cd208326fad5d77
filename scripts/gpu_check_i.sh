#!/bin/bash
mkdir -p gpurun_out
timeout 900 python scripts/lanes_probe.py > gpurun_out/r2b_lanes_probe.log 2>&1; tail -5 gpurun_out/r2b_lanes_probe.log
