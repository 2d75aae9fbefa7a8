#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "refill or all_51 or forced or special" 2>&1 | tail -5
for rf in 1 0; do
  NRLDPC_REFILL=$rf python bench.py --workload bg2_z52_r15_it8et_b65536 --steps 30 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('refill=$rf cfg3 stop', round(d['value'],3), round(d['ms_per_step'],4), d['config']['mean_iters'])"
done
python bench.py --workload bg2_z52_r15_it8_b65536 --steps 30 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 fixed', round(d['value'],3), round(d['ms_per_step'],4))"
for Z in 8 32 96 192; do for rf in 1 0; do
  NRLDPC_REFILL=$rf python tools/gpu_point.py --bg 1 --Z $Z --rate 1/3 --early-term --esn0 0.5 --reps 10 | sed "s/^/refill=$rf /"
done; done
