#!/bin/bash
# one full ncu capture of the float32 decode kernel with the parity-check stop (headline code), source page as CSV
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/prof_et \
  python bench.py --workload bg1_z384_r13_it8et_b4096 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt > gpurun_out/ncu_et.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_et.ncu-rep gpurun_out/prof_et.summary.txt > /dev/null
ncu -i gpurun_out/prof_et.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/prof_et.source.csv.gz
rm -f gpurun_out/prof_et.ncu-rep
ls -la gpurun_out | tail -5
