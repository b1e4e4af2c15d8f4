#!/bin/bash
# round 2, 1 GPU: GPU test-suite with the per-cell list expansion / register-tiled bitmask walk / split contact integration,
# launch list of one rebuild at n = 1e6, configs[2] line, configs[1] A/B of MIS_CONTACT_SPLIT
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2s_pytest_gpu.log
timeout 300 python scripts/profile_rebuild.py 1000000 > gpurun_out/r2s_plain_rebuild.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_launches_rebuild_n1m.csv python scripts/profile_rebuild.py 1000000 > gpurun_out/r2s_ncu_rebuild.log 2>&1
timeout 600 python bench.py --mode rebuild --steps 20 --warmup 5 --no-cpu > gpurun_out/r2s_bench_rebuild.json 2> gpurun_out/r2s_bench_rebuild.err
for sp in 0 1; do
MIS_CONTACT_SPLIT=$sp timeout 600 python bench.py --mode configs1 --steps 200 --warmup 20 --no-cpu > gpurun_out/r2s_configs1_split$sp.json 2> gpurun_out/r2s_configs1_split$sp.err
done
python - <<'PY'
import csv, json
try:
    rows=[r for r in csv.reader(open('gpurun_out/r2s_launches_rebuild_n1m.csv')) if len(r)>10]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
    data=rows[1:]; half=len(data)//2; tot={}
    for r in data[-half:]:
        k=r[ik].split('(')[0][:60]; tot[k]=tot.get(k,0)+float(r[iv].replace(',',''))
    for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:12]: print('%10.1f us  %s'%(v/1000, k))
    print('sum us', sum(tot.values())/1000, len(data))
except Exception as e: print('launch list ERR', e)
for f in ('r2s_bench_rebuild','r2s_configs1_split0','r2s_configs1_split1'):
    try:
        d=json.loads([l for l in open('gpurun_out/%s.json'%f) if l.startswith('{')][0])
        print(f, d['value'], d['ms_per_step'], d.get('rebuild_ms'), d.get('step_ms'), d.get('contact'), d.get('e2e',{}).get('value'))
    except Exception as e: print(f, 'ERR', e)
PY
tail -5 gpurun_out/r2s_pytest_gpu.log
