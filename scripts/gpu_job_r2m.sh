#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r2m_tile_ab.jsonl 2> gpurun_out/r2m_tile_ab.err
timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r2m_rebuild.json 2> gpurun_out/r2m_rebuild.err
python - <<'PY'
import json
for l in open('gpurun_out/r2m_tile_ab.jsonl'):
    d=json.loads(l)
    print(d['n'], {m: (round(d[m]['deform_us'],1), round(d[m]['force_us'],1), round(d[m]['step_us_chained'],1)) for m in ('mode0','mode1','mode2')})
d=json.load(open('gpurun_out/r2m_rebuild.json'))
print('rebuild', d['rebuild_ms'], 'step', d['step_ms'], d['value'])
PY
tail -3 gpurun_out/r2m_rebuild.err
