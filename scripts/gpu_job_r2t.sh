#!/bin/bash
# ncu --set full of the neighbour-build kernels at n = 1e6 (first build of the process)
mkdir -p gpurun_out
timeout 300 python scripts/profile_rebuild.py 1000000 > gpurun_out/r2t_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_tile_expand|k_tile_walk_bits' -c 3 -o gpurun_out/prof_r2t_build python scripts/profile_rebuild.py 1000000 > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log
