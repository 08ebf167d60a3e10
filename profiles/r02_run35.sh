#!/bin/bash
# round 2 (session 3), evidence part A: the whole gpu suite on the shipped library + the cost of the rigid-coupled path at config 3's size
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/r02f_gputests.log 2>&1; tail -4 gpurun_out/r02f_gputests.log
timeout 300 python profiles/rigid_cost.py --steps 100 > gpurun_out/r02f_rigid_cost.json 2> gpurun_out/r02f_rigid_cost.err; tail -c 1200 gpurun_out/r02f_rigid_cost.json; tail -3 gpurun_out/r02f_rigid_cost.err
