#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for wl in bg1_z384_r13_it8_b4096 bg1_z384_r13_it8et_b4096 bg1_z384_r89_it20et_b4096; do
  python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-e2e --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', round(d['value'],3), round(d['ms_per_step'],4), 'f16x2', round(d['f16x2']['value'],3), round(d['f16x2']['ms_per_step'],4))"
done
for wl in bg2_z52_r15_it8et_lowsnr_b65536 bg2_z52_r15_it8et_b65536; do for rf in 1 0; do
  NRLDPC_REFILL=$rf python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('refill=$rf $wl', round(d['value'],3), round(d['ms_per_step'],4), d['config']['mean_iters'], 'f16x2', round(d['f16x2']['value'],3))"
done; done
for Z in 4 8 16 32 52 96 192; do for rf in 1 0; do
  NRLDPC_REFILL=$rf python tools/gpu_point.py --bg 1 --Z $Z --rate 1/3 --early-term --esn0 0.5 --reps 10 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('refill=$rf Z',d['Z'],d['gbps'],d['ms'],round(d['mean_iters'],2))"
done; done
