#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_errors_and_edges.py tests/test_gpu_parity.py tests/test_gpu_gather_modes.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -6 gpurun_out/r2x_pytest.log
