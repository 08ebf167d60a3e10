#!/bin/bash
mkdir -p gpurun_out
L=$PWD/taichi_mpm_b200/lib
echo "== 1 GPU: validate build, chunk 1, 1500 substeps"; MPMB_LIB=$L/libmpmb_validate.so timeout 300 python profiles/repro_flow.py 1.0 1500 1 2>&1 | tail -1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "== 2 GPUs: plain, chunk 500, 6000 substeps"; timeout 300 $TR --master-port 29571 profiles/repro_flow2.py 1.0 6000 500 peer 2>&1 | grep "rank [01]:" | tail -2
echo "== 2 GPUs: bench --gpus 2"
timeout 900 $TR --master-port 29572 bench.py --gpus 2 > gpurun_out/r02_run18_bench2.json 2> gpurun_out/r02_run18_bench2.err
tail -c 4500 gpurun_out/r02_run18_bench2.json; grep -i "error\|rank" gpurun_out/r02_run18_bench2.err | head -5
