cd $GRAFT_REPO_ROOT
python scripts/pair_exp.py 100000 2>&1 | tail -1
python scripts/pair_exp.py 1000000 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout=300 2>&1 | tail -4
