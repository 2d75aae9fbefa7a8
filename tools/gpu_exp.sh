#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for dt in f32 f16x2; do
    python bench.py --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --llr-dtype $dt 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$dt',round(d['value'],3),round(d['ms_per_step'],4))"
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:decode_nms -s 3 -c 1 --csv --log-file gpurun_out/dram_$dt.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --llr-dtype $dt > /dev/null 2>&1
    tail -3 gpurun_out/dram_$dt.csv | cut -d, -f13-
done
