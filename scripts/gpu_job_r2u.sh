#!/bin/bash
# build kernels: list tests + launch list of one rebuild at n = 1e6
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gather_modes.py tests/test_gpu_parity.py tests/test_gpu_errors_and_edges.py -m gpu -q -x > gpurun_out/r2u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2u_pytest.log
timeout 300 python scripts/profile_rebuild.py 1000000 > gpurun_out/r2u_plain_rebuild.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none --csv --log-file gpurun_out/r2u_launches_rebuild_n1m.csv python scripts/profile_rebuild.py 1000000 > gpurun_out/r2u_ncu_rebuild.log 2>&1
python - <<'PY'
import csv
try:
    rows=[r for r in csv.reader(open('gpurun_out/r2u_launches_rebuild_n1m.csv')) if len(r)>10]
    hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); im=hdr.index('Metric Name'); iid=hdr.index('ID')
    data=rows[1:]
    ids=sorted({int(r[iid]) for r in data}); half=ids[len(ids)//2]
    tot={}
    for r in data:
        if int(r[iid])<half: continue
        k=r[ik].split('(')[0][:50]; d=tot.setdefault(k,{})
        v=float(r[iv].replace(',',''))
        if 'time' in r[im]: d['us']=d.get('us',0)+v/1000
        elif 'inst_executed.sum' in r[im]: d['inst']=d.get('inst',0)+v
        elif 'issue_active' in r[im]: d['issue']=v
        elif 'thread_inst' in r[im]: d['thr']=v
    for k,d in sorted(tot.items(), key=lambda kv:-kv[1].get('us',0))[:12]: print('%9.1f us  inst %.3g issue %.0f%% thr %.1f  %s'%(d.get('us',0), d.get('inst',0), d.get('issue',0), d.get('thr',0), k))
    print('sum us', sum(d.get('us',0) for d in tot.values()))
except Exception as e: print('launch list ERR', e)
PY
tail -4 gpurun_out/r2u_pytest.log
