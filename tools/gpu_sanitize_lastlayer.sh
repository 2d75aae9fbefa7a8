#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sanitizer_lastlayer.txt
for tool in memcheck racecheck synccheck; do
  for args in "0 1 384 3 1" "1 1 384 3 1"; do
    echo "== compute-sanitizer --tool $tool tools/gpu_repro.py $args" >> gpurun_out/sanitizer_lastlayer.txt
    timeout 200 compute-sanitizer --tool $tool python tools/gpu_repro.py $args 2>&1 | grep -E "hard equal|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -8 >> gpurun_out/sanitizer_lastlayer.txt
  done
done
cat gpurun_out/sanitizer_lastlayer.txt
