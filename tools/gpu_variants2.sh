#!/bin/bash
# A/B of library builds (NRLDPC_B200_LIB): headline fixed / stop and config 4, float32 and packed-half, two rounds each
b() {  # lib workload dtype
  NRLDPC_B200_LIB=$PWD/$1 python bench.py --workload $2 --steps 60 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $3 2>&1 | tail -1 |
    python -c "import json,sys;d=json.loads(sys.stdin.read());print('$(basename $1) $2 $3',round(d['value'],3),'Gb/s',round(d['ms_per_step'],4),'ms',flush=True)"
}
for round in 1 2; do
for lib in ldpc_3gpp_matlab_b200/libnrldpc_b200.so build_variants/*.so; do
  for dt in f32 f16x2; do
    b $lib bg1_z384_r13_it8_b4096 $dt
    b $lib bg1_z384_r13_it8et_b4096 $dt
    b $lib bg1_z384_r89_it20et_b4096 $dt
  done
done; done 2>&1 | tee gpurun_out/variants2.txt
