cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q --timeout=300 2>&1 | tail -3
bash scripts/job_multi.sh 2 n2_p2p --steps 200 --warmup 20 --particles 100000 --halo p2p
bash scripts/job_multi.sh 2 n2_nccl --steps 200 --warmup 20 --particles 100000 --halo nccl
