#!/bin/bash
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "all_51 or special or config4 or live_handles" 2>&1 | tail -4
python tools/gpu_shape_scan.py --bg 1 2>&1 | grep '^{' > gpurun_out/r02/shape_scan_bg1_f32.jsonl
python tools/gpu_shape_scan.py --bg 1 --dtype f16x2 --zs 36,48,72,104,128,144,192,208,224,256,288 2>&1 | grep '^{' > gpurun_out/r02/shape_scan_bg1_f16.jsonl
python tools/gpu_shape_scan.py --bg 2 --zs 36,52,104,128,144,208,288 2>&1 | grep '^{' > gpurun_out/r02/shape_scan_bg2_f32.jsonl
python - <<'PY'
import json
for f in ("bg1_f32","bg1_f16","bg2_f32"):
    rows=[json.loads(l) for l in open(f"gpurun_out/r02/shape_scan_{f}.jsonl")]
    print("==",f)
    for Z in sorted({r["Z"] for r in rows}):
        rr=[r for r in rows if r["Z"]==Z]
        print(Z, " ".join(f'{r["shape"]}:{r.get("gbps","ERR")}{"" if r.get("same_bits",True) else "!DIFF"}' for r in rr))
PY
for wl in bg1_z384_r89_it20et_b4096 bg1_z384_r13_it8_b4096; do python bench.py --workload $wl --steps 50 --no-cpu-baseline --no-e2e --no-side 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$wl', d['value'], d['ms_per_step'], d['f16x2']['value'])"; done
