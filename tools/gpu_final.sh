#!/bin/bash
# last visit of the round: the whole GPU suite on the committed tree, smoke, the default bench line, and the
# phase-lock experiment (parity-check stop armed but never taken)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" timeout 120 python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
for dt in f32 f16x2; do
  b stop_never_taken bg1_z384_r13_it8et_lowsnr_b4096 $dt X=1
  b fixed bg1_z384_r13_it8_b4096 $dt X=1
  b stop bg1_z384_r13_it8et_b4096 $dt X=1
done 2>&1 | tee gpurun_out/phase.txt
python bench.py 2>&1 | tail -1 > gpurun_out/bench.json; cut -c1-200 gpurun_out/bench.json
