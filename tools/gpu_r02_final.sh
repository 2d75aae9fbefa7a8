#!/bin/bash
# r02 final 1-GPU evidence visit: test-suite, smoke, default bench (+ reference arm), the other BASELINE configurations,
# ncu --set full + DRAM metric pass + launch list of the COMMITTED binary, latency, BLER-loop rate, sanitizer, sweep.
T=${1:-r02_v5}
O=gpurun_out/$T
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1 || tail -20 $O/build.log
timeout 900 python -m pytest tests -x -q -m gpu -rs 2>&1 | tail -8 | tee $O/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
python bench.py --steps 100 2> $O/bench.err | tail -1 > $O/bench_n1.json; tail -3 $O/bench.err
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference.json
for wl in bg1_z384_r13_it8et_b4096 bg2_z52_r15_it8_b65536 bg2_z52_r15_it8et_b65536 bg1_z384_r89_it20et_b4096 bg1_z384_r13_it8et_lowsnr_b4096; do
  python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > $O/bench_$wl.json
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/bench_*.json")):
    try: d=json.loads(open(f).read())
    except Exception as e: print(f,"ERR",e); continue
    print(f.split("/")[-1], round(d["value"],4),"Gb/s", round(d["ms_per_step"],4),"ms iters",d.get("config",{}).get("mean_iters"), "f16x2", (d.get("f16x2") or {}).get("value"), "e2e", (d.get("e2e") or {}).get("value"))
d=json.loads(open("$O/bench_n1.json").read())
print("f64_pageable", d["e2e"]["f64_pageable"]["value"], d["e2e"]["f64_pageable"]["ratio_to_pinned_f32"], "threads", d["e2e"]["f64_pageable"]["host_threads"])
print("config4", json.dumps(d["config4"]["points"]))
print("bler_loop", d["bler_loop"]["frames_per_s"], d["bler_loop"]["ms_per_batch"])
print("cpu", d["cpu_baseline"])
PY
for dt in f32 f16x2; do
  ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 3 -c 1 -f -o $O/decode_$dt \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt > $O/ncu_full_$dt.log 2>&1
  python tools/ncu_summary.py $O/decode_$dt.ncu-rep $O/decode_${dt}_ncu_full.txt > /dev/null
  rm -f $O/decode_$dt.ncu-rep
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:decode_nms -s 3 -c 1 --csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt 2>/dev/null | grep -E '"dram__|"gpu__time|^"ID"' > $O/dram_$dt.csv
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/gpu_latency.py > $O/latency.log 2>&1; cp gpurun_out/latency.json $O/latency.json
python tools/gpu_bler_rate.py > $O/bler_rate.log 2>&1; cp gpurun_out/bler_rate.json $O/bler_rate.json
: > $O/sanitizer.txt
for tool in memcheck racecheck synccheck; do
  for args in "0 1 384 3 1" "1 1 384 3 1" "0 2 52 15 1" "0 1 208 3 1" "1 1 208 3 1" "0 1 8 200 1" "0 2 8 300 1"; do
    echo "== compute-sanitizer --tool $tool tools/gpu_repro.py $args" >> $O/sanitizer.txt
    timeout 400 compute-sanitizer --tool $tool python tools/gpu_repro.py $args 2>&1 | grep -E "hard equal|ERROR SUMMARY|RACECHECK SUMMARY|Error|error" | head -6 >> $O/sanitizer.txt
  done
done
tail -30 $O/sanitizer.txt
python tools/sweep.py --mb 400 --out $O/sweep_1gpu > $O/sweep.log 2>&1; tail -3 $O/sweep.log
du -sh $O
