#!/bin/bash
# round 2, GPU call 3: full gpu suite, new bench.py (smoke at reduced scale, then the real line), ncu of the three tile kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02_run3_gputests.log 2>&1
tail -4 gpurun_out/r02_run3_gputests.log
timeout 300 python bench.py --scale 0.25 --steps 5 --warmup 3 --develop 50 --frames 1 --frame-substeps 5 --cpu-sample 20000 > gpurun_out/r02_run3_smoke.json 2> gpurun_out/r02_run3_smoke.err
tail -c 1500 gpurun_out/r02_run3_smoke.json; tail -5 gpurun_out/r02_run3_smoke.err
timeout 900 python bench.py > gpurun_out/r02_run3_bench.json 2> gpurun_out/r02_run3_bench.err
tail -c 3000 gpurun_out/r02_run3_bench.json; tail -5 gpurun_out/r02_run3_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_p2g|k_g2p|k_grid" -s 9 -c 3 -f -o gpurun_out/r02_run3_prof \
    python bench.py --steps 4 --warmup 3 --frames 0 --no-cpu-baseline --develop 0 > gpurun_out/r02_run3_ncu.log 2>&1
tail -2 gpurun_out/r02_run3_ncu.log
