cd $GRAFT_REPO_ROOT
nvidia-smi topo -m 2>&1 | head -8
timeout 600 python -m pytest tests/test_gpu_slab.py -m gpu -x -q --timeout=300 2>&1 | tail -8
for halo in p2p nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 200 --warmup 20 --halo $halo > gpurun_out/bench_n2_$halo.json 2> gpurun_out/bench_n2_$halo.err
  tail -c 400 gpurun_out/bench_n2_$halo.err
done
python - <<'PY'
import json
for f in ("p2p","nccl"):
    try:
        d=json.load(open(f"gpurun_out/bench_n2_{f}.json")); print(f, d["value"], d["ms_per_step"], d["steady_state"]["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["config"].get("halo"))
    except Exception as e: print(f, "ERR", e)
PY
