#!/bin/bash
# round 2, 2 GPUs: cross-process slab tests (fused P2P halo push and NCCL), then the default bench line at N = 2 (pair-work slab cuts)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02_pytest_2gpu.log
timeout 600 python -m pytest -m gpu tests/test_gpu_slab.py -q -rs >> gpurun_out/r02_pytest_2gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err
echo "bench rc=$?" >> gpurun_out/r2v_bench_n2.err
tail -4 gpurun_out/r02_pytest_2gpu.log; tail -c 300 gpurun_out/r2v_bench_n2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2v_bench_n2.json') if l.startswith('{')][0])
    print(d['value'], d['ms_per_step'], d['parity_check'], d.get('ms_per_step_by_rank'), d.get('gather_ms_by_rank'), d['config'].get('owned_per_gpu'), d.get('e2e',{}).get('value'))
except Exception as e: print('ERR', e)
PY
