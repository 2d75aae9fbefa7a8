#!/bin/bash
# DRAM traffic of one decode launch (headline batch) with / without the persisting access-policy window over the c2v scratch
mkdir -p gpurun_out/r02
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "forced or all_51 or config4" 2>&1 | tail -3
for win in 1 0; do for dt in f32 f16x2; do
  NRLDPC_L2_WINDOW=$win ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:decode_nms -s 3 -c 2 --csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt 2>/dev/null | grep -E "dram__|gpu__time" | awk -F'","' -v w=$win -v d=$dt '{print "window="w, d, $5, $(NF-2), $(NF-1), $NF}' | tr -d '"'
done; done | tee gpurun_out/r02/dram_window_ab.txt
for win in 1 0; do for dt in f32 f16x2; do
  NRLDPC_L2_WINDOW=$win python bench.py --steps 50 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('window=$win $dt', round(d['value'],3), round(d['ms_per_step'],4))"
  NRLDPC_L2_WINDOW=$win python bench.py --workload bg1_z384_r13_it8et_b4096 --steps 50 --no-cpu-baseline --no-e2e --no-alt --no-side --llr-dtype $dt 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('window=$win $dt stop', round(d['value'],3), round(d['ms_per_step'],4))"
done; done | tee -a gpurun_out/r02/dram_window_ab.txt
