#!/bin/bash
# A/B of the early-termination ('Parity check satisfied') syndrome variants: parity tests first, then the headline
# (fixed iterations: regression check), the headline / config 3 / config 4 with the parity-check stop under the
# default thresholds and with the staged / bit-sliced variants switched off or forced by environment.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
for dt in f32 f16x2; do
  b fixed bg1_z384_r13_it8_b4096 $dt X=1
  b default bg1_z384_r13_it8et_b4096 $dt X=1
  b unrolled_unstaged bg1_z384_r13_it8et_b4096 $dt NRLDPC_BITSLICED_MIN_ROWS=99 NRLDPC_STAGED_MIN_ROWS=99
  b unrolled_staged bg1_z384_r13_it8et_b4096 $dt NRLDPC_BITSLICED_MIN_ROWS=99
  b default bg1_z384_r89_it20et_b4096 $dt X=1
  b bitsliced bg1_z384_r89_it20et_b4096 $dt NRLDPC_BITSLICED_MIN_ROWS=4
  b default bg2_z52_r15_it8et_b65536 $dt X=1
  b unstaged bg2_z52_r15_it8et_b65536 $dt NRLDPC_STAGED_MIN_ROWS=99
done 2>&1 | tee gpurun_out/et_ab.txt
python tools/gpu_bler_rate.py 2>&1 | grep NMS | tee -a gpurun_out/et_ab.txt
