#!/bin/bash
# round 2 (session 3): P2G occupancy — both flush arenas overlaid on the row staging area (6.5 KB less shared memory) and a
# register cap (launch bound 5 CTAs/SM: 168 registers, no spills): 512-row chunks fit 5 CTAs per SM, 576-row chunks stay at 4
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py "MPMB_EXP_P2G_OVERLAY+MPMB_P2G_CH=512+MPMB_P2G_MINB=5:p2g_5cta_ch512" "MPMB_EXP_P2G_OVERLAY+MPMB_P2G_MINB=5:p2g_168reg_ch576" "MPMB_EXP_P2G_OVERLAY+MPMB_P2G_CH=448+MPMB_P2G_MINB=5:p2g_5cta_ch448" --reps 2 --steps 200 > gpurun_out/r02_ab_p2g_occupancy.log 2>&1
grep -v "^{" gpurun_out/r02_ab_p2g_occupancy.log | cut -c1-420
