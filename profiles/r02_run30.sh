#!/bin/bash
# round 2 (session 3): A/B of the P2G cell->thread balancing variants (bank-pair swap, population rank) against the default
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py MPMB_EXP_P2G_PAIR:pair MPMB_EXP_P2G_RANK:rank --reps 2 --steps 200 > gpurun_out/r02_ab_p2g_balance.log 2>&1
tail -8 gpurun_out/r02_ab_p2g_balance.log | cut -c1-600
