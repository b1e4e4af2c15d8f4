#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/debug_grad.py > gpurun_out/r2i_debug_grad.log 2>&1
cat gpurun_out/r2i_debug_grad.log | tail -50
