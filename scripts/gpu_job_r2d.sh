#!/bin/bash
# ncu --set full of the shared-memory tile kernels (one step at n = 1e5)
mkdir -p gpurun_out
export MIS_GATHER=1
python scripts/profile_step.py 100000 4 > gpurun_out/plain_tile.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_deform_t|k_force_t|k_deform_fin' -s 3 -c 3 -o gpurun_out/prof_r02_tile python scripts/profile_step.py 100000 4 > gpurun_out/ncu_tile.log 2>&1
tail -3 gpurun_out/ncu_tile.log; cat gpurun_out/plain_tile.log
