#!/bin/bash
# after a change to the decode kernels' code generation: the other configurations and a spot of the sweep
for wl in bg2_z52_r15_it8_b65536 bg2_z52_r15_it8et_b65536 bg1_z384_r89_it20et_b4096 bg1_z384_r13_it8_b4096 bg1_z384_r13_it8et_b4096; do
  python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-e2e --no-side 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$wl',round(d['value'],3),round(d['ms_per_step'],4),'f16x2',round(d['f16x2']['value'],3))"
done
python tools/sweep.py --zs 8,52,96,208,384 --mb 400 --out gpurun_out/sweep_check > /dev/null 2>&1; grep -E "^\| (8|52|96|208|384) " gpurun_out/sweep_check.md
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_51 or golden or special or refill or forced" 2>&1 | tail -2
