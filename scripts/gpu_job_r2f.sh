#!/bin/bash
# round 2, 8 GPUs: the default bench line (10 M body, strong scaling, obstacle in contact) and configs[3] (64 scenes, 8 per GPU)
mkdir -p gpurun_out
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r2f_bench_n$N.json 2> gpurun_out/r2f_bench_n$N.err
echo "bench rc=$?" >> gpurun_out/r2f_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --mode batch --steps 200 --warmup 10 > gpurun_out/r2f_batch_n$N.json 2> gpurun_out/r2f_batch_n$N.err
echo "batch rc=$?" >> gpurun_out/r2f_batch_n$N.err
tail -c 300 gpurun_out/r2f_bench_n$N.err; head -c 400 gpurun_out/r2f_bench_n$N.json; echo; tail -c 300 gpurun_out/r2f_batch_n$N.err; head -c 400 gpurun_out/r2f_batch_n$N.json
