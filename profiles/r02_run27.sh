#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29621 bench.py --gpus 2 > gpurun_out/r02_run27_bench2.json 2> gpurun_out/r02_run27_bench2.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run27_bench2.json").read().strip().splitlines()[-1])
print({k:(round(v["ms_per_step"],4), round(v["value"]), {a:round(b,4) for a,b in v["stage_ms_per_step"].items()}) for k,v in d["states"].items() if isinstance(v,dict)}, "launches", d["gpu_launches"], "alive", d["alive_particles"], "parity", d["parity_check"]["ok"], d["parity_check"]["max_err"], "weak", round(d["weak_scaling"]["value"]), round(d["weak_scaling"]["ms_per_step"],4), "e2e", round(d["e2e"]["value"]))
PY
grep -i "error" gpurun_out/r02_run27_bench2.err | head -3
