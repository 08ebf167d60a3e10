#!/bin/bash
mkdir -p gpurun_out
for g in 1 0; do
  echo "== bench MPMB_GRAPH=$g"
  MPMB_GRAPH=$g timeout 600 python bench.py --no-cpu-baseline --frames 0 > gpurun_out/r02_run25_bench_g$g.json 2> gpurun_out/r02_run25_bench_g$g.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02_run25_bench_g$g.json").read().strip().splitlines()[-1])
print({k:(round(v["ms_per_step"],4), {a:round(b,4) for a,b in v["stage_ms_per_step"].items()}) for k,v in d["states"].items() if isinstance(v,dict)}, "launches", d["gpu_launches"], "alive", d["alive_particles"])
PY
done
