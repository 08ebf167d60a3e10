#!/bin/bash
# smoke of the snow / water workloads on the slab path at reduced scale (2 GPUs) before the expensive full-size runs
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for w in snow256:0.5 water512:0.25; do
  timeout 300 $TR --master-port $((29600 + ${#w})) bench.py --gpus 2 --workload ${w%%:*} --scale ${w##*:} --steps 20 --warmup 5 --develop 300 --frames 1 --frame-substeps 100 --also-weak 0 > gpurun_out/r02_run20_${w%%:*}.json 2> gpurun_out/r02_run20_${w%%:*}.err
  echo "== $w"; tail -c 1800 gpurun_out/r02_run20_${w%%:*}.json; grep -i "error" gpurun_out/r02_run20_${w%%:*}.err | head -3
done
