# usage: gpu_job_ab.sh VAR   -- gather-kernel timings with VAR=0 and default, then the parity tests
cd $GRAFT_REPO_ROOT
VAR=${1:-MIS_SPLIT_LISTS}
env $VAR=0 python scripts/pair_exp.py 100000 2>&1 | tail -1
python scripts/pair_exp.py 100000 2>&1 | tail -1
env $VAR=0 python scripts/pair_exp.py 1000000 2>&1 | tail -1
python scripts/pair_exp.py 1000000 2>&1 | tail -1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors_and_edges.py -m gpu -x -q --timeout=300 2>&1 | tail -4
