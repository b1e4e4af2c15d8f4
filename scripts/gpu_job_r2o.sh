#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2o_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -5 gpurun_out/r2o_pytest.log
timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r2o_rebuild.json 2> gpurun_out/r2o_rebuild.err
MIS_BUILD_WALK=1 timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r2o_rebuild_walk.json 2> gpurun_out/r2o_rebuild_walk.err
timeout 600 python scripts/tile_ab.py 100000 > gpurun_out/r2o_tile_ab.jsonl 2> gpurun_out/r2o_tile_ab.err
python - <<'PY'
import json
for f in ('r2o_rebuild','r2o_rebuild_walk'):
    d=json.load(open(f'gpurun_out/{f}.json')); print(f, 'rebuild', d['rebuild_ms'], 'step', d['step_ms'], d['value'])
for l in open('gpurun_out/r2o_tile_ab.jsonl'):
    d=json.loads(l)
    print(d['n'], {m: (round(d[m]['deform_us'],1), round(d[m]['force_us'],1), round(d[m]['step_us_chained'],1)) for m in ('mode0','mode1','mode2')})
PY
