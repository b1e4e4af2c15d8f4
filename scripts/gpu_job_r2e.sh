#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/tile_ab.py 100000 1000000 > gpurun_out/r2e_tile_ab.jsonl 2> gpurun_out/r2e_tile_ab.err
echo "ab rc=$?" >> gpurun_out/r2e_tile_ab.err
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
timeout 600 python bench.py --mode configs1 --steps 100 --warmup 10 --no-cpu > gpurun_out/r2e_bench_configs1.json 2> gpurun_out/r2e_bench_configs1.err
python - <<'PY'
import json
for l in open('gpurun_out/r2e_tile_ab.jsonl'):
    d=json.loads(l)
    print(d['n'], {m: (round(d[m]['deform_us'],1), round(d[m]['force_us'],1), round(d[m]['step_us_chained'],1)) for m in ('mode0','mode1','mode2')}, d['dx_between_modes'], d['dv_between_modes'])
d=json.load(open('gpurun_out/r2e_bench_configs1.json'))
print(d['value'], d['ms_per_step'], d['contact'], d['steady_state'], d['roofline']['kernels'])
PY
tail -3 gpurun_out/r2e_tile_ab.err; tail -4 gpurun_out/r2e_pytest.log
