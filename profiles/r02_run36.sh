#!/bin/bash
# round 2 (session 3): rigid-coupled path cost after the row permutation in k_p2g_rigid
mkdir -p gpurun_out
timeout 300 python profiles/rigid_cost.py --steps 100 > gpurun_out/r02g_rigid_cost.json 2> gpurun_out/r02g_rigid_cost.err; tail -c 1200 gpurun_out/r02g_rigid_cost.json; tail -3 gpurun_out/r02g_rigid_cost.err
