#!/bin/bash
# instruction-cache counters of the float32 decode kernel, fixed iterations vs parity-check stop, 2 and 1 CTA per SM
mkdir -p gpurun_out
ncu --query-metrics 2>/dev/null | grep -i -E "^(sm__icc|gcc__|idc__|smsp__inst_executed_pipe|sm__inst_fetch|lts__t_.*srcunit_(ltcfabric|gcc))" | cut -c1-120 > gpurun_out/icache_metrics_available.txt
M=$(grep -E "^(sm__icc_requests|gcc__requests|gcc__)" gpurun_out/icache_metrics_available.txt | awk '{print $1}' | grep -E "icc_requests|gcc__requests|gcc__.*(hit|miss|lookup)" | head -24 | paste -sd, -)
echo "metrics: $M" > gpurun_out/icache.txt
for wl in bg1_z384_r13_it8_b4096 bg1_z384_r13_it8et_b4096; do
  for cap in 296 148; do
    echo "== $wl grid cap $cap" >> gpurun_out/icache.txt
    NRLDPC_GRID_CAP=$cap ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,lts__t_bytes.sum,$M \
      --clock-control none -k regex:decode_nms -s 3 -c 1 python bench.py --workload $wl --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt 2>&1 |
      grep -E "^\s+(gpu__|smsp__|lts__|sm__icc|gcc__)" >> gpurun_out/icache.txt
  done
done
cat gpurun_out/icache.txt | cut -c1-150
