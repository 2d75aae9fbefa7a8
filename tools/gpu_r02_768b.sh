#!/bin/bash
for lib in ldpc_3gpp_matlab_b200/libnrldpc_b200.so build_variants/lib768.so; do
 for wl in bg2_z52_r15_it8et_b65536 bg2_z52_r15_it8_b65536 bg2_z52_r15_it8et_lowsnr_b65536; do
  NRLDPC_SHAPE_MODEL=0 NRLDPC_B200_LIB=$PWD/$lib python bench.py --workload $wl --steps 30 --no-cpu-baseline --no-e2e --no-side --no-alt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$(basename $lib) $wl', round(d['value'],3), round(d['ms_per_step'],4), d['config']['mean_iters'])"
 done
done
