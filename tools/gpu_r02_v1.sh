#!/bin/bash
# r02 visit 2 (1 GPU): the GPU test-suite (incl. MEX gateway, config 4, pageable host path), smoke, the default bench line
# with its new side keys, the reference arm.
mkdir -p gpurun_out/r02
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r02/build.log 2>&1 || tail -20 gpurun_out/r02/build.log
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/r02/smoke.log
python bench.py --steps 50 2> gpurun_out/r02/bench.err | tail -1 > gpurun_out/r02/bench_n1.json; tail -5 gpurun_out/r02/bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02/bench_n1.json").read())
print("value",d["value"],"ms",d["ms_per_step"],"e2e",d["e2e"]["value"],"f16tr",d["e2e"]["f16_transport"]["value"],"f64pageable",d["e2e"]["f64_pageable"])
print("config4",json.dumps(d["config4"]["points"]))
print("bler_loop",json.dumps(d["bler_loop"]))
print("cpu",d["cpu_baseline"])
PY
python tools/gpu_latency.py > gpurun_out/r02/latency.log 2>&1; tail -12 gpurun_out/r02/latency.log
