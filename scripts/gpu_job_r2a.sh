#!/bin/bash
# round 2, first GPU pass: parity tests, smoke (incl. the MLP contact path), the new bench line at N = 1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --deselect tests/test_gpu_golden.py::test_trajectories_vs_reference_source > gpurun_out/r2a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2a_bench_n1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
tail -3 gpurun_out/r2a_pytest.log; tail -2 gpurun_out/r2a_smoke.log; tail -c 600 gpurun_out/r2a_bench_n1.err
