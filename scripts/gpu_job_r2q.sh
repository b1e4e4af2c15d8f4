#!/bin/bash
mkdir -p gpurun_out
for fp in 0 1; do
MIS_FORCE_PERSIST=$fp timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r2q_tile_ab_fp$fp.jsonl 2> gpurun_out/r2q_tile_ab_fp$fp.err
done

python - <<'PY'
import json
for fp in (0,1):
    for l in open(f'gpurun_out/r2q_tile_ab_fp{fp}.jsonl'):
        d=json.loads(l)
        print('persist',fp, d['n'], {m: (round(d[m]['deform_us'],1), round(d[m]['force_us'],1), round(d[m]['step_us_chained'],1)) for m in ('mode0','mode2')}, d['dx_between_modes'])
PY

