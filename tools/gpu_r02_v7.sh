#!/bin/bash
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/gpu_latency.py > gpurun_out/r02/latency.log 2>&1; grep -E "batch': 1,|kernel only|RX chain|gateway|ZERO" gpurun_out/r02/latency.log
bash tools/gpu_r02_768b.sh
