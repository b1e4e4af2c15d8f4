#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_parity_scene.py 2.0 > gpurun_out/r2k_debug.log 2>&1
grep -c "finite True" gpurun_out/r2k_debug.log; grep "finite False" gpurun_out/r2k_debug.log | head -3; grep "step 120" gpurun_out/r2k_debug.log
