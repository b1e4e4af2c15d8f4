cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_slab.py -m gpu -x -q --timeout=300 2>&1 | tail -4
bash scripts/job_multi.sh 2 n2_strong_b --steps 100 --warmup 10 --mode strong --n-total 10000000
bash scripts/job_multi.sh 2 n2_weak_b --steps 100 --warmup 10
