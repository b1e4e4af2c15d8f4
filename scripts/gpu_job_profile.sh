cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | tail -6
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -c 300 gpurun_out/bench_j.err
python scripts/profile_contact.py 100000 6 > gpurun_out/plain_j.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 80 --csv --log-file gpurun_out/launches_j.csv python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_j1.log 2>&1
python scripts/profile_contact.py 100000 6 > gpurun_out/plain_j.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_deform_c|k_force_c|k_sdf_chain_sk' -s 164 -c 4 -o gpurun_out/prof_r01_j python scripts/profile_contact.py 100000 6 > gpurun_out/ncu_j2.log 2>&1
tail -3 gpurun_out/ncu_j2.log
timeout 900 python scripts/run_configs.py 0 2 3 > gpurun_out/configs_j.jsonl 2> gpurun_out/configs_j.err; cat gpurun_out/configs_j.jsonl; tail -c 500 gpurun_out/configs_j.err
