#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
for dt in f32 f16x2; do
  b corelooped bg1_z384_r13_it8_b4096 $dt X=1
  b corelooped bg1_z384_r13_it8et_b4096 $dt X=1
  b corelooped bg1_z384_r89_it20et_b4096 $dt X=1
done 2>&1 | tee gpurun_out/ab3.txt
