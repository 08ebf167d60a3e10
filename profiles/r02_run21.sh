#!/bin/bash
# round 2, GPU call 21 (8 GPUs): strong scaling of config 3 at N=8, BASELINE config 5 (water 512^3, 64 M, 8 GPUs), config 4 (snow, 16 M, 4 GPUs)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 420 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 > gpurun_out/r02_bench_sand256_8gpu.json 2> gpurun_out/r02_bench_sand256_8gpu.err
tail -c 1500 gpurun_out/r02_bench_sand256_8gpu.json; grep -i "error" gpurun_out/r02_bench_sand256_8gpu.err | head -3
timeout 600 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --workload water512 --also-weak 0 > gpurun_out/r02_bench_water512_8gpu.json 2> gpurun_out/r02_bench_water512_8gpu.err
tail -c 1500 gpurun_out/r02_bench_water512_8gpu.json; grep -i "error" gpurun_out/r02_bench_water512_8gpu.err | head -3
timeout 420 $TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --workload snow256 --also-weak 0 > gpurun_out/r02_bench_snow256_4gpu.json 2> gpurun_out/r02_bench_snow256_4gpu.err
tail -c 1500 gpurun_out/r02_bench_snow256_4gpu.json; grep -i "error" gpurun_out/r02_bench_snow256_4gpu.err | head -3
