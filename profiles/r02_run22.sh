#!/bin/bash
# round 2, GPU call 22: full gpu suite + 1-GPU bench with the P2G excess path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_run22_gputests.log 2>&1; tail -3 gpurun_out/r02_run22_gputests.log
timeout 900 python bench.py > gpurun_out/r02_run22_bench.json 2> gpurun_out/r02_run22_bench.err; tail -c 2500 gpurun_out/r02_run22_bench.json; tail -3 gpurun_out/r02_run22_bench.err
