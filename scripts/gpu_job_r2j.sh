#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r2j_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r2j_bench_n1.json 2> gpurun_out/r2j_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r2j_bench_n1.err
tail -8 gpurun_out/r2j_pytest.log; tail -3 gpurun_out/r2j_smoke.log; tail -c 300 gpurun_out/r2j_bench_n1.err; head -c 300 gpurun_out/r2j_bench_n1.json
