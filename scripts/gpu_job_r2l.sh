#!/bin/bash
# N GPUs: default bench line only
mkdir -p gpurun_out
N=${1:-4}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2l_bench_n$N.json 2> gpurun_out/r2l_bench_n$N.err
echo "bench rc=$?" >> gpurun_out/r2l_bench_n$N.err
tail -c 300 gpurun_out/r2l_bench_n$N.err; head -c 300 gpurun_out/r2l_bench_n$N.json
