#!/bin/bash
# round 2, 1 GPU: the whole GPU test-suite, smoke, the default bench line, configs[2] (--mode rebuild), gather-mode A/B
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
echo "bench rc=$?" >> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_rebuild.json 2> gpurun_out/r02_bench_rebuild.err
timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r02_tile_ab.jsonl 2> gpurun_out/r02_tile_ab.err
tail -6 gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_smoke.log; tail -c 200 gpurun_out/r02_bench_n1.err; head -c 250 gpurun_out/r02_bench_n1.json; echo; head -c 400 gpurun_out/r02_bench_rebuild.json
