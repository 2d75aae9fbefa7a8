#!/bin/bash
# compute-sanitizer over the early-termination (two-stage syndrome) paths only: one codeword per CTA (bit-sliced) and
# several codewords per CTA (per-thread, staged), float32 and packed half; then the looped-variant comparison.
mkdir -p gpurun_out
: > gpurun_out/sanitizer_et.txt
for tool in memcheck racecheck synccheck; do
  for args in "0 1 384 3 1" "1 1 384 3 1" "0 2 52 15 1" "1 2 52 15 1" "0 1 224 3 1"; do
    echo "== compute-sanitizer --tool $tool tools/gpu_repro.py $args" >> gpurun_out/sanitizer_et.txt
    timeout 300 compute-sanitizer --tool $tool python tools/gpu_repro.py $args 2>&1 | grep -E "hard equal|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -8 >> gpurun_out/sanitizer_et.txt
  done
done
cat gpurun_out/sanitizer_et.txt
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
b loop bg1_z384_r13_it8_b4096 f32 NRLDPC_DECODE_VARIANT=loop | tee gpurun_out/loop_variant.txt
b loop bg1_z384_r13_it8et_b4096 f32 NRLDPC_DECODE_VARIANT=loop | tee -a gpurun_out/loop_variant.txt
