#!/bin/bash
mkdir -p gpurun_out/r02
export NRLDPC_SHAPE_MODEL=0
for bg in 1 2; do for dt in f32 f16x2; do
  python tools/gpu_shape_scan.py --bg $bg --dtype $dt --zs all 2>&1 | grep '^{' > gpurun_out/r02/shape_full_bg${bg}_${dt}.jsonl
  wc -l gpurun_out/r02/shape_full_bg${bg}_${dt}.jsonl
done; done
