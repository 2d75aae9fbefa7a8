#!/bin/bash
# compute-sanitizer (memcheck + racecheck + synccheck) over small decodes of both arithmetic modes and the chain kernels.
mkdir -p gpurun_out
: > gpurun_out/sanitizer.txt
for tool in memcheck racecheck synccheck; do
  for args in "0 1 384 3 1" "1 1 384 3 1" "0 2 52 15 0" "1 2 6 131 1" "2 1 384 2 1" "2 2 52 9 0"; do
    echo "== compute-sanitizer --tool $tool tools/gpu_repro.py $args" >> gpurun_out/sanitizer.txt
    timeout 600 compute-sanitizer --tool $tool python tools/gpu_repro.py $args 2>&1 | grep -E "hard equal|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -8 >> gpurun_out/sanitizer.txt
  done
done
echo "== compute-sanitizer --tool memcheck pytest -k 'rate or encode or modulate or crc'" >> gpurun_out/sanitizer.txt
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -x -q -m gpu -k "rate or encode or modulate or crc or chain" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | tail -3 >> gpurun_out/sanitizer.txt
cat gpurun_out/sanitizer.txt
