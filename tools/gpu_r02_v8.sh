#!/bin/bash
# r02 verification visit: GPU suite, racecheck / synccheck of the shapes that changed (one-codeword CTAs with a partly filled
# last warp, per-slot filter flag of the multi-codeword kernels under the stop, refill kernel), spot rows of the sweep.
T=${1:-r02_v6}
O=gpurun_out/$T
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -8 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
: > $O/sanitizer.txt
for tool in racecheck synccheck memcheck; do
  for args in "0 2 52 15 1" "0 1 208 3 1" "1 1 208 3 1" "0 1 8 200 1" "0 2 8 300 1" "0 1 384 3 1" "1 2 52 15 1" "0 1 208 3 0"; do
    echo "== compute-sanitizer --tool $tool tools/gpu_repro.py $args" >> $O/sanitizer.txt
    timeout 400 compute-sanitizer --tool $tool python tools/gpu_repro.py $args 2>&1 | grep -E "hard equal|ERROR SUMMARY|RACECHECK SUMMARY|Error|error|Warning" | head -6 >> $O/sanitizer.txt
  done
done
cat $O/sanitizer.txt
python tools/sweep.py --zs 144,208,224,288,384 --mb 400 --out $O/sweep_spot > $O/sweep_spot.log 2>&1; grep -E "^\| (144|208|224|288|384) " $O/sweep_spot.md
python bench.py --steps 50 --no-cpu-baseline --no-e2e --no-alt --no-side 2>/dev/null | tail -1 > $O/bench_quick.json; python -c "
import json; d=json.load(open('$O/bench_quick.json')); print('headline', d['value'], d['ms_per_step'])"
