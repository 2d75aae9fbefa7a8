#!/bin/bash
# Quick GPU visit: parity tests + short bench + full ncu capture of the decode kernel.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 > gpurun_out/bench_quick.json
python -c "import json;d=json.load(open('gpurun_out/bench_quick.json'));print(d['value'],d['ms_per_step'],d['config']['bler_at_esn0'],d['clocks'])"
ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/prof_decode python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full.log 2>&1
