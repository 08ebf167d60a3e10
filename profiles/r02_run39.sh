#!/bin/bash
# round 2 (session 3): the whole gpu suite again on the final rigid kernels (per-thread impulses, epoch marks) with the new tests:
# CPIC edge cases, the reference's solver object with rigid bodies through libmpmb, config 2 at full size with two bodies
mkdir -p gpurun_out
(time timeout 900 python -m pytest tests -x -q -m gpu) > gpurun_out/r02i_gputests.log 2>&1; tail -4 gpurun_out/r02i_gputests.log
timeout 300 python -m pytest tests/test_gpu_zz_full_parity.py -q -m gpu -s -k rigid 2>&1 | grep -E "config 2|passed|failed" | cut -c1-300 > gpurun_out/r02i_full_size_rigid.log; cat gpurun_out/r02i_full_size_rigid.log
