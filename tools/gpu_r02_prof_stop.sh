#!/bin/bash
# ncu --set full + source page of the multi-codeword kernel under the stop (config 3) and the headline stop path
mkdir -p gpurun_out/r02
cap() { # name workload
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/r02/$1 \
    python bench.py --workload $2 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side > gpurun_out/r02/$1.log 2>&1
  python tools/ncu_summary.py gpurun_out/r02/$1.ncu-rep gpurun_out/r02/$1.summary.txt > /dev/null
  ncu -i gpurun_out/r02/$1.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/r02/$1.source.csv.gz
  rm -f gpurun_out/r02/$1.ncu-rep
}
cap cfg3_stop bg2_z52_r15_it8et_b65536
cap cfg3_fixed bg2_z52_r15_it8_b65536
cap headline_stop bg1_z384_r13_it8et_b4096
cap headline_fixed bg1_z384_r13_it8_b4096
ls -la gpurun_out/r02 | tail -12
