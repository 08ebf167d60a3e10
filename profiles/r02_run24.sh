#!/bin/bash
mkdir -p gpurun_out
timeout 900 python profiles/ab_variants.py "MPMB_GRID_MINB=6:grid6" "MPMB_GRID_MINB=8:grid8" --reps 1 --steps 200 > gpurun_out/r02_run24_ab.log 2>&1; grep -E "^default|^MPMB" gpurun_out/r02_run24_ab.log | cut -c1-700
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02_run24_bench.json 2> gpurun_out/r02_run24_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_run24_bench.json").read().strip().splitlines()[-1])
print("e2e", d["e2e"]["value"], d["e2e"]["frame_seconds_all"], "value", d["value"], {k:(v["ms_per_step"] if v else None) for k,v in d["states"].items() if isinstance(v,dict)})
PY
