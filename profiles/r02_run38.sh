#!/bin/bash
# round 2 (session 3): ncu --set full of the rigid kernels (one substep's four launches after warm-up) at config 3's size
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_rigid|k_g2p_rigid|k_gather_cdf|k_cdf_raster" -s 80 -c 4 -f -o gpurun_out/r02f_prof_rigid \
    python profiles/rigid_cost.py --steps 3 > gpurun_out/r02f_ncu_rigid.log 2>&1; tail -2 gpurun_out/r02f_ncu_rigid.log | cut -c1-200
ls -la gpurun_out | grep rigid
