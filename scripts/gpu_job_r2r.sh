#!/bin/bash
# 8 GPUs: the default bench line with two settings of the obstacle ranks' load-balance cost
mkdir -p gpurun_out
N=${1:-8}
for cost in 100000; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 30 --warmup 5 --contact-cost $cost > gpurun_out/r2r_bench_n${N}_c$cost.json 2> gpurun_out/r2r_bench_n${N}_c$cost.err
echo "bench rc=$?" >> gpurun_out/r2r_bench_n${N}_c$cost.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2r_bench_n*_c*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][0])
        print(f, d['value'], d['ms_per_step'], d['parity_check']['ok'], d['ms_per_step_by_rank'], d['gather_ms_by_rank'])
    except Exception as e:
        print(f, 'ERR', e)
PY
