#!/bin/bash
# round 2, GPU call 4 (2 GPUs): strong scaling of config 3 on the fused peer exchange, with the parity check; flow statistics
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 \
    > gpurun_out/r02_run4_bench2.json 2> gpurun_out/r02_run4_bench2.err
tail -c 3500 gpurun_out/r02_run4_bench2.json; tail -5 gpurun_out/r02_run4_bench2.err
timeout 300 python profiles/flow_stats.py --scale 1.0 > gpurun_out/r02_run4_flow_stats.jsonl 2>&1
cat gpurun_out/r02_run4_flow_stats.jsonl
timeout 300 python -m pytest tests/test_gpu_slab.py -x -q -m gpu 2>&1 | tail -3
