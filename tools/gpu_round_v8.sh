#!/bin/bash
# r01 v8 evidence visit (no large ncu reports): parity tests, smoke, bench lines of every BASELINE config (+ reference
# arm, + the headline with the parity-check stop), BLER-loop rate, small-call latency, ncu launch list of the bench command.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json
python bench.py --workload bg2_z52_r15_it8_b65536 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg3.json
python bench.py --workload bg1_z384_r89_it20et_b4096 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg4.json
python bench.py --workload bg1_z384_r13_it8et_b4096 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_headline_stop.json
python tools/gpu_bler_rate.py > gpurun_out/bler_rate.log 2>&1
python tools/gpu_latency.py > gpurun_out/latency.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
du -sh gpurun_out
