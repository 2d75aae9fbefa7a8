#!/bin/bash
# Quick GPU visit: short bench of both arithmetic modes + full ncu capture of each decode kernel.
mkdir -p gpurun_out
for dt in f32 f16x2; do
  python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --llr-dtype $dt 2>&1 | tail -1 > gpurun_out/bench_quick_$dt.json
  python -c "import json;d=json.load(open('gpurun_out/bench_quick_$dt.json'));print('$dt',d['value'],d['ms_per_step'],d['config']['bler_at_esn0'],d['clocks'])"
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/prof_decode_$dt python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --llr-dtype $dt > gpurun_out/ncu_full_$dt.log 2>&1
done
