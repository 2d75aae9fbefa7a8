#!/bin/bash
# chain-kernel visit: CRC (table steps), random-bits kernel, vectorised block-error count: tests, BLER-loop rate, launch list
O=gpurun_out/r02_bler2; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bler.py tests/test_mex_gateway.py -x -q -m gpu 2>&1 | tail -5
python tools/gpu_bler_rate.py > $O/bler_rate.log 2>&1; cp gpurun_out/bler_rate.json $O/; grep "layered" $O/bler_rate.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/bler_launches.csv python tools/gpu_bler_prof.py > $O/bler_prof.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_bler.py -x -q -m gpu -k "random_bits or device_crc" 2>&1 | grep -E "passed|failed|ERROR SUMMARY" | head -3
