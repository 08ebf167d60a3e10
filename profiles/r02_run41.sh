#!/bin/bash
# round 2 (session 3): the gpu suite once more on the final tree (adds the mirror's AsyncMPM scheduler on the engine, seeded-lattice + rigid)
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -x -q -m gpu) > gpurun_out/r02k_gputests.log 2>&1; tail -5 gpurun_out/r02k_gputests.log
