#!/bin/bash
# Kernel experiments: the same benches against several builds of the library (NRLDPC_B200_LIB, build_variants/*.so).
mkdir -p gpurun_out
b() {  # label, workload, dtype, env...
  local label=$1 wl=$2 dt=$3; shift 3
  env "$@" python bench.py --workload $wl --steps 50 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$label $wl $dt',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms iters',d['config']['mean_iters'],flush=True)"
}
for lib in ldpc_3gpp_matlab_b200/libnrldpc_b200.so build_variants/*.so; do
  for dt in f32 f16x2; do
    b $(basename $lib) bg1_z384_r13_it8et_b4096 $dt NRLDPC_B200_LIB=$PWD/$lib
    b $(basename $lib) bg1_z384_r13_it8_b4096 $dt NRLDPC_B200_LIB=$PWD/$lib
  done
done 2>&1 | tee gpurun_out/variants.txt
