#!/bin/bash
# round 2, 2 GPUs: the cross-process slab tests (fused P2P halo push and NCCL) with 0 skipped, then the bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_pytest_2gpu.log
python -m pytest -m gpu tests/test_gpu_slab.py -q -rs >> gpurun_out/r02_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_2gpu.log
MIS_GATHER=1 python -m pytest -m gpu tests/test_gpu_slab.py -q -rs >> gpurun_out/r02_pytest_2gpu.log 2>&1
echo "pytest (MIS_GATHER=1) rc=$?" >> gpurun_out/r02_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2c_bench_n2.err
tail -6 gpurun_out/r02_pytest_2gpu.log; tail -c 400 gpurun_out/r2c_bench_n2.err; head -c 600 gpurun_out/r2c_bench_n2.json
