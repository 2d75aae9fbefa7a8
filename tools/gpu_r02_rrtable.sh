#!/bin/bash
# gather table of rate recovery: parity tests, A/B timing (NRLDPC_RR_TABLE=0/1), launch list of a BLER batch
O=gpurun_out/r02_rrtable; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bler.py -x -q -m gpu -k "rate_match or harq or full_chain or fused or bler or round_trip" 2>&1 | tail -4
for t in 0 1; do
  NRLDPC_RR_TABLE=$t python tools/gpu_bler_rate.py 2>&1 | grep layered | sed "s/^/rr_table=$t /"
done | tee $O/ab.txt
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bler_launches.csv python tools/gpu_bler_prof.py > $O/bler_prof.log 2>&1
python - <<PY
import csv
rows=list(csv.reader(open("$O/bler_launches.csv")))
hdr=None; seq=[]
for r in rows:
    if r and r[0]=='ID': hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        if d['Metric Name']=='gpu__time_duration.sum':
            v=float(d['Metric Value'].replace(',','')); u=d['Metric Unit']; v*= {'ns':1e-3,'us':1,'ms':1e3}.get(u,1)
            seq.append((d['Kernel Name'][:60],v))
idx=max(i for i,(n,v) in enumerate(seq) if 'random_bits' in n)
for n,v in seq[idx:]: print("%-62s %8.1f"%(n,v))
PY
python tools/gpu_latency.py 2>&1 | grep -E "RX chain"
