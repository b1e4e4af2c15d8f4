#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fp64_and_grad.py -m gpu -q -x > gpurun_out/r2h_pytest_fp64.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2h_pytest_fp64.log
tail -30 gpurun_out/r2h_pytest_fp64.log
