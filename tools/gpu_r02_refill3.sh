#!/bin/bash
# prefetched-refill kernel (NRLDPC_REFILL=3): parity, sanitizer, then A/B against the group kernel / old refill kernel
O=gpurun_out/r02_refill3; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "refill" 2>&1 | tail -5
for tool in racecheck memcheck; do
  NRLDPC_REFILL=3 timeout 300 compute-sanitizer --tool $tool python tools/gpu_repro.py 0 2 52 300 1 2>&1 | grep -E "hard equal|SUMMARY|Error" | head -5
  NRLDPC_REFILL=3 timeout 300 compute-sanitizer --tool $tool python tools/gpu_repro.py 0 1 8 900 1 2>&1 | grep -E "hard equal|SUMMARY|Error" | head -5
done
for wl in bg2_z52_r15_it8et_b65536 bg2_z52_r15_it8et_lowsnr_b65536; do
  for cfg in "0 2" "1 2" "3 1" "3 2" "3 3" "3 4"; do set -- $cfg
    NRLDPC_REFILL=$1 NRLDPC_REFILL_SPARES=$2 python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('refill=$1 spares=$2 $wl', round(d['value'],3), round(d['ms_per_step'],4), d['config']['mean_iters'])"
  done
done | tee $O/cfg3.txt
for bg in 1 2; do for Z in 4 8 16 24 32 52 64 96 128 192; do
  for cfg in "0 2" "1 2" "3 1" "3 2" "3 3"; do set -- $cfg
    NRLDPC_REFILL=$1 NRLDPC_REFILL_SPARES=$2 python tools/gpu_point.py --bg $bg --Z $Z --rate 1/3 --early-term --esn0 0.5 --reps 10 --mb 200 | sed "s/^/refill=$1 spares=$2 /"
  done
done; done | tee $O/sweep.txt
