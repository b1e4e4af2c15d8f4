cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q --timeout=400 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -c 300 gpurun_out/bench_j.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 300 gpurun_out/bench_ref.err; head -c 700 gpurun_out/bench_ref.json; echo
bash scripts/gpu_job_multi.sh 1 n1_1250k --steps 100 --warmup 10 --n 1250000 --no-obstacle --no-cpu
