cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_deepsdf.py -m gpu -x -q --timeout=120 2>&1 | tail -4
python scripts/chain_time.py 2>&1 | tail -12
python scripts/contact_exp.py 2>&1 | tail -1
