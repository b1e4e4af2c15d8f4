#!/bin/bash
mkdir -p gpurun_out
python scripts/profile_rebuild.py 1000000 > gpurun_out/plain_rebuild.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_rebuild_n1m.csv python scripts/profile_rebuild.py 1000000 > gpurun_out/ncu_rebuild.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r02_launches_rebuild_n1m.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
data=rows[1:]
half=len(data)//2
tot={}
for r in data[-half:]:
    k=r[ik].split('(')[0][:60]; tot[k]=tot.get(k,0)+float(r[iv].replace(',',''))
for k,v in sorted(tot.items(), key=lambda kv:-kv[1])[:20]: print('%10.1f us  %s'%(v/1000 if v>5000 else v, k))
print(len(data))
PY
