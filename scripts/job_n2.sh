cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_n${N}_big.json 2> gpurun_out/bench_n${N}_big.err

grep -v "^\*\|OMP_NUM\|^$\|Elapsed\|Maximum\|^\s" gpurun_out/bench_n${N}_big.err | tail -5
tail -1 gpurun_out/bench_n${N}_big.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['n_gpus'], '%.3e' % d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], '%.3e' % d['e2e']['value'], d['config']['n_particles'], d['config']['owned_per_gpu'], d['config']['ghosts_per_gpu'], d['roofline']['kernels'])
"
