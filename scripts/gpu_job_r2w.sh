#!/bin/bash
# ncu --set full (with source) of the two gather kernels of a large scene (n = 1e6: k_deform_t + the persistent k_force_p)
mkdir -p gpurun_out
timeout 300 python scripts/profile_step.py 1000000 4 > gpurun_out/r2w_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k 'regex:k_force_p|k_deform_t' -s 2 -c 2 -o gpurun_out/prof_r2w_step1m python scripts/profile_step.py 1000000 4 > gpurun_out/r2w_ncu.log 2>&1
tail -3 gpurun_out/r2w_ncu.log
