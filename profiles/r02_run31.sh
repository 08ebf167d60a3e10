#!/bin/bash
# round 2 (session 3): new materials on the hardware + A/B of k_g2p before/after the three new material branches
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_reference_golden.py -x -q -m gpu -k "single_substep or multi_step_other or reference_substeps or reference_transfers" > gpurun_out/r02_materials_gputests.log 2>&1; tail -3 gpurun_out/r02_materials_gputests.log
for rep in 1 2; do for L in prev new; do
  if [ $L = prev ]; then export MPMB_LIB=$PWD/taichi_mpm_b200/lib/libmpmb_prev.so; else unset MPMB_LIB; fi
  echo "== $L $rep" >> gpurun_out/r02_ab_materials_g2p.log
  timeout 300 python bench.py --steps 200 --warmup 20 --frames 0 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); st=l['states']
print({k:(round(v['ms_per_step'],4), {a:round(b,4) for a,b in v['stage_ms_per_step'].items()}) for k,v in st.items() if isinstance(v,dict)})" >> gpurun_out/r02_ab_materials_g2p.log 2>&1
done; done
cat gpurun_out/r02_ab_materials_g2p.log | cut -c1-400
