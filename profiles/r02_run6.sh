#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python profiles/repro_flow.py 1.0 20000 5000 > gpurun_out/r02_run6_big.log 2>&1; tail -3 gpurun_out/r02_run6_big.log
timeout 300 python profiles/flow_stats.py --scale 1.0 > gpurun_out/r02_run6_stats.log 2>&1; tail -3 gpurun_out/r02_run6_stats.log | cut -c1-600
timeout 900 compute-sanitizer --tool memcheck --print-limit 8 python profiles/flow_stats.py --scale 0.5 --develop 2000 > gpurun_out/r02_run6_memcheck.log 2>&1; grep -v "^{" gpurun_out/r02_run6_memcheck.log | head -60
