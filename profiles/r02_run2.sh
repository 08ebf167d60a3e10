#!/bin/bash
# round 2, GPU call 2: full gpu suite on the TMA/FFMA2 engine, A/B of FFMA2 and G2P occupancy, launch list
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_run2_gputests.log 2>&1
tail -5 gpurun_out/r02_run2_gputests.log
timeout 1200 python profiles/ab_variants.py "MPMB_G2P_MINB=4:minb4" "MPMB_EXP_SCALAR_F4:scalar_f4" "MPMB_EXP_SCALAR_F4+MPMB_G2P_MINB=4:scalar_f4_minb4" --reps 1 > gpurun_out/r02_run2_ab.log 2>&1
tail -8 gpurun_out/r02_run2_ab.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_run2_launches.csv \
    python bench.py --steps 3 --warmup 3 --frames 0 --no-cpu-baseline > gpurun_out/r02_run2_launch.log 2>&1
tail -2 gpurun_out/r02_run2_launch.log
