#!/bin/bash
# r02 visit (1 GPU): GPU test-suite, smoke, headline bench (fixed + stop), cfg3, latency
mkdir -p gpurun_out/r02
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 50 --no-cpu-baseline 2> gpurun_out/r02/bench.err | tail -1 > gpurun_out/r02/bench_n1.json; tail -3 gpurun_out/r02/bench.err
for wl in bg1_z384_r13_it8et_b4096 bg2_z52_r15_it8_b65536 bg2_z52_r15_it8et_b65536 bg1_z384_r89_it20et_b4096; do
  python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 > gpurun_out/r02/bench_$wl.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02/bench_*.json")):
    try: d=json.loads(open(f).read())
    except Exception as e: print(f,"ERR",e); continue
    print(f.split("/")[-1], round(d["value"],3),"Gb/s", round(d["ms_per_step"],4),"ms iters",d["config"]["mean_iters"], "f16x2", d.get("f16x2",{}).get("value"), d.get("f16x2",{}).get("ms_per_step"))
PY
python tools/gpu_latency.py > gpurun_out/r02/latency.log 2>&1; grep -E "batch': 1,|kernel only|RX chain" gpurun_out/r02/latency.log
