#!/bin/bash
# GPU visit for the sum-product (reference algorithm) mode: its parity tests, then the whole GPU suite, then the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "bp or sum_product or reference_algorithm or decode64" 2>&1 | tail -15 > gpurun_out/pytest_bp.log; cat gpurun_out/pytest_bp.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 2>&1 | tail -1 | tee gpurun_out/bench_bp.json | python -c "import json,sys;d=json.loads(sys.stdin.read());print(d['value'],d['ms_per_step'],d.get('reference_algorithm_on_gpu'),d['cpu_baseline'])"
