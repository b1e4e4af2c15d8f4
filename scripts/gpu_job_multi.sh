# usage: job_multi.sh N tag [bench args...]   -- one multi-GPU bench run, result line into gpurun_out/bench_<tag>.json
cd $GRAFT_REPO_ROOT
N=$1; TAG=$2; shift 2
if [ "$N" = "1" ]; then
  timeout 1500 python bench.py --gpus 1 "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
fi
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_$TAG.err | tail -5
tail -1 gpurun_out/bench_$TAG.json | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
c = d['config']
print(d['n_gpus'], d['scaling'], '%.4e' % d['value'], 'ms/step', round(d['ms_per_step'], 4), 'steady', round(d['steady_state']['ms_per_step'], 4), 'e2e %.3e' % d['e2e']['value'], 'n', c['n_particles'], c.get('owned_per_gpu'), c.get('ghosts_per_gpu'), {k: round(v['ms'], 4) for k, v in d['roofline']['kernels'].items()}, d['clocks'])
"
