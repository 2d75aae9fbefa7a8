#!/bin/bash
# Multi-GPU visit: bench at N = 1, 2, 4, 8 (whatever the box has), then (unless SKIP_SWEEP is set) the config-5 sweep on all GPUs.
# SKIP_N1=1 skips the single-GPU run (taken from the 1-GPU visit).
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "GPUs: $NG"
[ -n "$SKIP_N1" ] || python bench.py --gpus 1 --steps 100 --warmup 3 2>&1 | tail -1 > gpurun_out/scale_n1.json
for n in 2 4 8; do
  if [ $n -le $NG ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 100 --warmup 3 2>&1 | grep '^{"metric"' | tail -1 > gpurun_out/scale_n$n.json
  fi
done
for f in gpurun_out/scale_n*.json; do python -c "import json,sys;d=json.load(open('$f'));print(d['n_gpus'],round(d['value'],3),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value'],3),'f16t',round(d['e2e']['f16_transport']['value'],3),'numa',d['e2e']['numa_bound'],'h2',round(d['f16x2']['value'],3),'bp',round(d['reference_algorithm_on_gpu']['value'],3),d['clocks']['sm_mhz'],d['clocks']['reasons'])"; done
[ -n "$SKIP_SWEEP" ] || python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/sweep.py --out gpurun_out/sweep_${NG}gpu > gpurun_out/sweep_${NG}gpu.log 2>&1
[ -n "$SKIP_SWEEP" ] || tail -3 gpurun_out/sweep_${NG}gpu.log
