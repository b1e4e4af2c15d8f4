#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/debug_parity_scene.py 12.8 > gpurun_out/r2g_debug.log 2>&1
for cf in 1 0; do
MIS_CONTACT_FIRST=$cf timeout 300 python bench.py --mode configs1 --steps 100 --warmup 10 --no-cpu > gpurun_out/r2g_configs1_cf$cf.json 2> gpurun_out/r2g_configs1_cf$cf.err
done
python - <<'PY'
import json
for cf in (1,0):
    d=json.load(open(f'gpurun_out/r2g_configs1_cf{cf}.json'))
    print('contact_first',cf, d['value'], d['ms_per_step'], d['steady_state']['ms_per_step'], d['contact']['in_contact_band_avg'])
PY
grep -n "finite False" -B3 -A2 gpurun_out/r2g_debug.log | head -40; grep -c "finite True" gpurun_out/r2g_debug.log; head -3 gpurun_out/r2g_debug.log
