cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_deepsdf.py -m gpu -x -q --timeout=120 2>&1 | tail -8
run() { env "$@" timeout 300 python scripts/contact_exp.py 2>&1 | tail -1; }
run MIS_SDF_CHAIN=0 MIS_SERIAL_CONTACT=1
run MIS_SERIAL_CONTACT=1
run A=1
run MIS_SDF_L2_PIN=0
timeout 300 python scripts/contact_time.py 2>&1 | tail -14
