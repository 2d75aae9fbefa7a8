#!/bin/bash
# CTA-shape scan, second part: narrow CTAs (at most 96 threads) at 12 / 16 / 24 resident CTAs per SM
O=gpurun_out/r02_shapes_hi; mkdir -p $O
export NRLDPC_SHAPE_MODEL=0
ZS=2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,18,20,22,24,26,28,30,32,36,40,44,48,52,56,60,64,72,80,88,96
for dt in f32 f16x2; do
  python tools/gpu_shape_scan.py --bg 1 --dtype $dt --zs $ZS --early-term --esn0 0.6 --only-extra --extra-caps 12,16,24 2>&1 | grep '^{' > $O/shape_stop_bg1_$dt.jsonl
  python tools/gpu_shape_scan.py --bg 2 --dtype $dt --zs $ZS --early-term --esn0 1.2 --only-extra --extra-caps 12,16,24 2>&1 | grep '^{' > $O/shape_stop_bg2_$dt.jsonl
  python tools/gpu_shape_scan.py --bg 1 --dtype $dt --zs $ZS --only-extra --extra-caps 12,16,24 2>&1 | grep '^{' > $O/shape_full_bg1_$dt.jsonl
  python tools/gpu_shape_scan.py --bg 2 --dtype $dt --zs $ZS --only-extra --extra-caps 12,16,24 2>&1 | grep '^{' > $O/shape_full_bg2_$dt.jsonl
done
wc -l $O/*.jsonl
grep '"Z": 52,' $O/shape_stop_bg2_f32.jsonl
