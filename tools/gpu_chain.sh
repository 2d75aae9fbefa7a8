#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12 --csv --log-file gpurun_out/chain_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-alt > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/chain_launches.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
hdr=rows[hi]; k=hdr.index('Kernel Name'); m=hdr.index('Metric Name'); v=hdr.index('Metric Value'); idc=hdr.index('ID')
d={}
for r in rows[hi+1:]:
    if len(r)>v: d.setdefault((r[idc],r[k][:48]),{})[r[m]]=float(r[v].replace(',',''))
for (i,n),x in d.items():
    t=x.get('gpu__time_duration.sum',0); b=x.get('dram__bytes_read.sum',0)+x.get('dram__bytes_write.sum',0)
    print(f"{n:50s} {t/1e3:9.1f} us  dram {b/1e6:8.1f} MB  {b/max(t,1):6.2f} GB/s" )
PY
