#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r2p_tile_ab.jsonl 2> gpurun_out/r2p_tile_ab.err
timeout 900 python -m pytest tests/test_gpu_gather_modes.py tests/test_gpu_parity.py -m gpu -q > gpurun_out/r2p_pytest.log 2>&1
python - <<'PY'
import json
for l in open('gpurun_out/r2p_tile_ab.jsonl'):
    d=json.loads(l)
    print(d['n'], {m: (round(d[m]['deform_us'],1), round(d[m]['force_us'],1), round(d[m]['step_us_chained'],1)) for m in ('mode0','mode1','mode2')}, d['dx_between_modes'])
PY
tail -5 gpurun_out/r2p_pytest.log
