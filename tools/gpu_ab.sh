#!/bin/bash
# A/B helper: parity tests, then the headline bench twice per arithmetic mode and the early-termination config once.
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for dt in f32 f16x2; do for i in 1 2; do
python bench.py --steps 100 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$dt',round(d['value'],3),round(d['ms_per_step'],4))"
done; done
for wl in bg1_z384_r89_it20et_b4096 bg2_z52_r15_it8_b65536; do for dt in f32 f16x2; do
python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$wl $dt',round(d['value'],3),round(d['ms_per_step'],4),d['config']['mean_iters'])"
done; done
