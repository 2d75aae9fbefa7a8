#!/bin/bash
for wl in bg2_z52_r15_it8et_lowsnr_b65536 bg2_z52_r15_it8et_b65536; do for rf in 1 0; do
  NRLDPC_REFILL=$rf python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('refill=$rf $wl', round(d['value'],3), round(d['ms_per_step'],4), d['config']['mean_iters'], d['config']['iters_hist'])"
done; done
python bench.py --workload bg2_z52_r15_it8_b65536 --steps 30 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 fixed', round(d['value'],3), round(d['ms_per_step'],4))"
