#!/bin/bash
# A/B the decode kernel variants on the headline workload (+ parity tests on the default variant).
for v in loop unroll; do
  echo "== variant $v"
  NRLDPC_DECODE_VARIANT=$v python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'],d['config']['bler_at_esn0'])"
done
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
NRLDPC_DECODE_VARIANT=loop python -m pytest tests -x -q -m gpu -k "decode" 2>&1 | tail -3
