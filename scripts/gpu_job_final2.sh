#!/bin/bash
# round 2 final, 1 GPU: the whole GPU test-suite, smoke, the default bench line, configs[2] (--mode rebuild), launch list and
# ncu --set full of the configs[1] step (with k_integrate behind the contact chain), ncu --set full of the final build kernels
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
( time timeout 900 python bench.py --steps 30 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2> gpurun_out/r02_bench_n1.time
echo "bench rc=$?" >> gpurun_out/r02_bench_n1.err
timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r02_bench_rebuild.json 2> gpurun_out/r02_bench_rebuild.err
python scripts/profile_contact.py 100000 6 > gpurun_out/plain_p2.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_n1.csv python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_p2a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_deform_t|k_deform_fin|k_force_c|k_integrate|k_sdf_chain_sk' -s 366 -c 6 -o gpurun_out/prof_r02_step python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_p2b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_tile_expand|k_tile_walk_bits' -c 3 -o gpurun_out/prof_r02_build python scripts/profile_rebuild.py 1000000 > gpurun_out/ncu_p2d.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu.log; tail -3 gpurun_out/r02_smoke.log; cat gpurun_out/r02_bench_n1.time; tail -c 200 gpurun_out/r02_bench_n1.err; head -c 300 gpurun_out/r02_bench_n1.json; echo; head -c 600 gpurun_out/r02_bench_rebuild.json; echo; tail -2 gpurun_out/ncu_p2a.log gpurun_out/ncu_p2b.log gpurun_out/ncu_p2d.log
