#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" timeout 120 python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
for dt in ${DTYPES:-f32}; do
  b refill bg2_z52_r15_it8et_b65536 $dt X=1
  b refill bg2_z52_r15_it8_b65536 $dt X=1
  b refill bg1_z384_r13_it8_b4096 $dt X=1
  b refill bg1_z384_r13_it8et_b4096 $dt X=1
done 2>&1 | tee gpurun_out/ab4.txt
timeout 200 python tools/gpu_bler_rate.py 2>&1 | grep NMS | tee -a gpurun_out/ab4.txt
