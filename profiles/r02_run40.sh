#!/bin/bash
# round 2 (session 3): the gpu suite on the final library (adds: delta_t changes, the reference's AsyncMPM scheduler on libmpmb)
mkdir -p gpurun_out
(time timeout 600 python -m pytest tests -x -q -m gpu) > gpurun_out/r02j_gputests.log 2>&1; tail -5 gpurun_out/r02j_gputests.log
