cd $GRAFT_REPO_ROOT
run() { env "$@" python scripts/contact_exp.py 2>&1 | tail -1; }
run MIS_SERIAL_CONTACT=1 MIS_SDF_L2_PIN=0
run MIS_SERIAL_CONTACT=1
run MIS_SDF_L2_PIN=0
run A=1
run MIS_DEFORM_CARVEOUT=50 MIS_SK_CARVEOUT=50
run MIS_DEFORM_CARVEOUT=40 MIS_SK_CARVEOUT=40
run MIS_DEFORM_CARVEOUT=100 MIS_SK_CARVEOUT=100
run MIS_DEFORM_CARVEOUT=50
run MIS_SERIAL_CONTACT=1 MIS_DEFORM_CARVEOUT=50
run MIS_SERIAL_CONTACT=1 MIS_DEFORM_CARVEOUT=100
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_deepsdf.py -m gpu -x -q --timeout=300 2>&1 | tail -5
