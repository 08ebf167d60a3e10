#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py "MPMB_P2G_CH=512:ch512" "MPMB_EXP_P2G_EXCESS:excess" --reps 1 --steps 200 > gpurun_out/r02_run23_ab.log 2>&1; tail -5 gpurun_out/r02_run23_ab.log | cut -c1-900
