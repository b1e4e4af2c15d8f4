#!/bin/bash
# round 2: shared-memory cell-tile gather kernels, A/B against the cluster kernels + parity in tile mode
mkdir -p gpurun_out
timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r2b_tile_ab.jsonl 2> gpurun_out/r2b_tile_ab.err
echo "ab rc=$?" >> gpurun_out/r2b_tile_ab.err
MIS_GATHER=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_golden.py tests/test_gpu_errors_and_edges.py tests/test_gpu_slab.py tests/test_gpu_streaming.py -m gpu -q --deselect tests/test_gpu_golden.py::test_trajectories_vs_reference_source > gpurun_out/r2b_pytest_tile.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2b_pytest_tile.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2b_smoke.log
cat gpurun_out/r2b_tile_ab.jsonl; tail -3 gpurun_out/r2b_tile_ab.err; tail -5 gpurun_out/r2b_pytest_tile.log; tail -3 gpurun_out/r2b_smoke.log
