#!/bin/bash
# round 2 profiles: launch list + ncu --set full of the default step (gather mode 2) with the contact chain, and k_sdf_gemm at 16384 rows
mkdir -p gpurun_out
python scripts/profile_contact.py 100000 6 > gpurun_out/plain_p2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_launches_n1.csv python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_p2a.log 2>&1
python scripts/profile_contact.py 100000 6 > gpurun_out/plain_p2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_deform_t|k_deform_fin|k_force_c|k_sdf_chain_sk' -s 305 -c 5 -o gpurun_out/prof_r02_step python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_p2b.log 2>&1
python scripts/sdf_bench.py 16384 > gpurun_out/plain_p2c.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:k_sdf_gemm$' -s 25 -c 1 -o gpurun_out/prof_r02_gemm python scripts/sdf_bench.py 16384 > gpurun_out/ncu_p2c.log 2>&1
tail -2 gpurun_out/ncu_p2a.log gpurun_out/ncu_p2b.log gpurun_out/ncu_p2c.log; cat gpurun_out/plain_p2.log | tail -2; grep -c k_ gpurun_out/r02_launches_n1.csv
