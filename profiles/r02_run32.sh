#!/bin/bash
# round 2 (session 3): CPIC rigid-coupled path on the hardware (parity tests) + compute-sanitizer racecheck/memcheck of one test + headline timing with the rigid hooks compiled in
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rigid.py -x -q -m gpu > gpurun_out/r02_rigid_gputests.log 2>&1; tail -5 gpurun_out/r02_rigid_gputests.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_rigid.py -x -q -m gpu -k "two_bodies" > gpurun_out/r02_rigid_memcheck.log 2>&1; tail -4 gpurun_out/r02_rigid_memcheck.log
timeout 300 python bench.py --steps 200 --warmup 20 --frames 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=l['states']
print({k:(round(v['ms_per_step'],4), {a:round(b,4) for a,b in v['stage_ms_per_step'].items()}) for k,v in st.items() if isinstance(v,dict)})" > gpurun_out/r02_bench_after_rigid.log 2>&1
cat gpurun_out/r02_bench_after_rigid.log | cut -c1-400
