#!/bin/bash
# GPU visit without the large ncu reports: parity tests, smoke, bench lines of every BASELINE config (+ reference arm),
# small-call latency, the ncu launch list of the bench command, and the config-5 sweep on one GPU.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json
python bench.py --workload bg2_z52_r15_it8_b65536 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg3.json
python bench.py --workload bg1_z384_r89_it20et_b4096 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg4.json
python tools/gpu_latency.py > gpurun_out/latency.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/sweep.py --out gpurun_out/sweep > gpurun_out/sweep.log 2>&1; tail -2 gpurun_out/sweep.log
du -sh gpurun_out
