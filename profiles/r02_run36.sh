#!/bin/bash
# round 2 (session 3): rigid-coupled path cost after per-thread impulse accumulation and epoch marks
mkdir -p gpurun_out
timeout 300 python profiles/rigid_cost.py --steps 100 > gpurun_out/r02h_rigid_cost.json 2> gpurun_out/r02h_rigid_cost.err; tail -c 1200 gpurun_out/r02h_rigid_cost.json; tail -3 gpurun_out/r02h_rigid_cost.err
