#!/bin/bash
# round 2 (session 3), evidence part B: the bench line of the shipped library, the reference arm, ncu --set full of the three tile
# kernels (the instantiations the bench times) and of the rigid kernels, the launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > gpurun_out/r02f_clocks.csv &
SMI=$!
timeout 900 python bench.py > gpurun_out/r02f_bench_final_1gpu.json 2> gpurun_out/r02f_bench_final_1gpu.err; tail -c 700 gpurun_out/r02f_bench_final_1gpu.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02f_bench_reference_arm.json 2> gpurun_out/r02f_bench_reference_arm.err; tail -c 500 gpurun_out/r02f_bench_reference_arm.json
kill $SMI
MPMB_GRAPH=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_g2p|k_grid" -s 9 -c 3 -f -o gpurun_out/r02f_prof \
    python bench.py --steps 4 --warmup 3 --frames 0 --no-cpu-baseline --develop 0 > gpurun_out/r02f_ncu.log 2>&1; tail -1 gpurun_out/r02f_ncu.log | cut -c1-200
MPMB_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02f_ncu_launches.csv \
    python bench.py --steps 3 --warmup 3 --frames 0 --no-cpu-baseline --develop 0 > gpurun_out/r02f_launch.log 2>&1; tail -1 gpurun_out/r02f_launch.log | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g_rigid|k_g2p_rigid|k_gather_cdf|k_cdf_raster" -s 92 -c 4 -f -o gpurun_out/r02f_prof_rigid \
    python profiles/rigid_cost.py --steps 3 > gpurun_out/r02f_ncu_rigid.log 2>&1; tail -2 gpurun_out/r02f_ncu_rigid.log | cut -c1-200
ls -la gpurun_out | tail -12
