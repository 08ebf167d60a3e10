#!/bin/bash
# round 2, GPU call 1: new parity tests, A/B of the round-1 kernel experiments, ncu of the product as timed
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r02_run1_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests/test_gpu_zz_full_parity.py -x -q -m gpu -s > gpurun_out/r02_run1_parity.log 2>&1
tail -5 gpurun_out/r02_run1_parity.log
timeout 1200 python profiles/ab_variants.py TILE_XYZ DUAL_ARENA SDF_FLAGS P2G_IPLANE TILE_XYZ+SDF_FLAGS+P2G_IPLANE TILE_XYZ+SDF_FLAGS+DUAL_ARENA --reps 1 > gpurun_out/r02_run1_ab.log 2>&1
tail -12 gpurun_out/r02_run1_ab.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_g2p|k_grid" -s 9 -c 3 -f -o gpurun_out/r02_run1_prof \
    python bench.py --steps 4 --warmup 3 --frames 0 --no-cpu-baseline > gpurun_out/r02_run1_ncu.log 2>&1
tail -3 gpurun_out/r02_run1_ncu.log
kill $SMI
