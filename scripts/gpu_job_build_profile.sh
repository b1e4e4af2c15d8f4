cd $GRAFT_REPO_ROOT
python scripts/profile_build.py > gpurun_out/plain_build.log 2>&1 && \
ncu --set full --clock-control none -k 'regex:rs_histogram|rs_scatter|scan_|k_cell_|k_pair_cells|k_apply_order|k_reintegrate|k_export_vec3|k_gather_vec3' -s 40 -c 40 -o gpurun_out/prof_r01_build python scripts/profile_build.py > gpurun_out/ncu_build.log 2>&1
tail -2 gpurun_out/ncu_build.log
python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -c 200 gpurun_out/bench_k.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_k.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['steady_state'])"
