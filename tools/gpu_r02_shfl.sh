#!/bin/bash
mkdir -p gpurun_out/r02_shfl
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "warp_shuffle" 2>&1 | tail -15
for tool in racecheck synccheck memcheck; do
  NRLDPC_DECODE_VARIANT=shfl timeout 300 compute-sanitizer --tool $tool python tools/gpu_repro.py 0 2 6 40 1 2>&1 | grep -E "hard equal|SUMMARY|Error" | head -5
  NRLDPC_DECODE_VARIANT=shfl timeout 300 compute-sanitizer --tool $tool python tools/gpu_repro.py 0 1 30 7 0 2>&1 | grep -E "hard equal|SUMMARY|Error" | head -5
done 2>&1 | tee gpurun_out/r02_shfl/sanitizer.txt
python tools/gpu_shfl_ab.py 2>&1 | tail -20
