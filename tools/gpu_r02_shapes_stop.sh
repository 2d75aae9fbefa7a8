#!/bin/bash
mkdir -p gpurun_out/r02
export NRLDPC_SHAPE_MODEL=0
python tools/gpu_shape_scan.py --bg 1 --dtype f32 --zs all --early-term --esn0 0.6 2>&1 | grep '^{' > gpurun_out/r02/shape_stop_bg1_f32.jsonl
python tools/gpu_shape_scan.py --bg 2 --dtype f32 --zs all --early-term --esn0 1.2 2>&1 | grep '^{' > gpurun_out/r02/shape_stop_bg2_f32.jsonl
python tools/gpu_shape_scan.py --bg 1 --dtype f16x2 --zs all --early-term --esn0 0.6 2>&1 | grep '^{' > gpurun_out/r02/shape_stop_bg1_f16x2.jsonl
python tools/gpu_shape_scan.py --bg 2 --dtype f16x2 --zs all --early-term --esn0 1.2 2>&1 | grep '^{' > gpurun_out/r02/shape_stop_bg2_f16x2.jsonl
wc -l gpurun_out/r02/shape_stop_*.jsonl
