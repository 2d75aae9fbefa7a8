#!/bin/bash
# One full GPU visit: parity tests, smoke, headline bench (+ reference arm), the other BASELINE configs, the ncu launch
# list of the bench command and one full ncu capture of each decode kernel, chain-kernel timings, small-call latency.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py 2>&1 | tail -1 | tee gpurun_out/bench.json | cut -c1-400
python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.json
python bench.py --workload bg2_z52_r15_it8_b65536 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg3.json
python bench.py --workload bg1_z384_r89_it20et_b4096 --steps 50 --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_cfg4.json
python tools/gpu_latency.py > gpurun_out/latency.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
for dt in f32 f16x2; do
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o gpurun_out/prof_decode_$dt python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt > gpurun_out/ncu_full_$dt.log 2>&1
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:decode_nms -s 3 -c 1 --csv --log-file gpurun_out/dram_$dt.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --llr-dtype $dt > /dev/null 2>&1
done
BP_STEPS=0 ncu --set full --clock-control none -k regex:decode_bp -c 1 -f -o gpurun_out/prof_bp2 python tools/gpu_bp_time.py 296 > gpurun_out/ncu_bp2.log 2>&1
ncu --set full --clock-control none -k regex:"encode_kernel|rate_match|rate_recover|qpsk" -c 4 -f -o gpurun_out/prof_chain python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-alt > gpurun_out/ncu_chain.log 2>&1
# gpurun copies back at most 64 MiB: keep the text summaries, drop the reports that carry the source pages
for r in gpurun_out/prof_*.ncu-rep; do python tools/ncu_summary.py $r > ${r%.ncu-rep}.summary.txt 2>&1; done
du -sm gpurun_out | awk '$1 > 60 {exit 1}' || rm -f gpurun_out/prof_decode_*.ncu-rep
ls -la gpurun_out
