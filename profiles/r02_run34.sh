#!/bin/bash
# round 2 (session 3): P2G accumulate loop software-pipelined (next particle's shared-memory row fetched one iteration ahead)
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py "MPMB_EXP_P2G_PREFETCH:p2g_prefetch" --reps 2 --steps 200 > gpurun_out/r02_ab_p2g_prefetch.log 2>&1
grep -v "^{" gpurun_out/r02_ab_p2g_prefetch.log | cut -c1-420
