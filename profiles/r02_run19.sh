#!/bin/bash
# round 2, GPU call 19 (4 GPUs): strong scaling of config 3 at N=4, and BASELINE config 4 (snow, 256^3, 16 M particles, 4 GPUs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29581 bench.py --gpus 4 > gpurun_out/r02_bench_sand256_4gpu.json 2> gpurun_out/r02_bench_sand256_4gpu.err
tail -c 3000 gpurun_out/r02_bench_sand256_4gpu.json; grep -i "error" gpurun_out/r02_bench_sand256_4gpu.err | head -5
timeout 900 $TR --master-port 29582 bench.py --gpus 4 --workload snow256 --also-weak 0 > gpurun_out/r02_bench_snow256_4gpu.json 2> gpurun_out/r02_bench_snow256_4gpu.err
tail -c 3000 gpurun_out/r02_bench_snow256_4gpu.json; grep -i "error" gpurun_out/r02_bench_snow256_4gpu.err | head -5
