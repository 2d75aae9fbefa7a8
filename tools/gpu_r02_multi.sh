#!/bin/bash
# r02 multi-GPU visit: the default bench line (with config4 and bler_loop side keys) under torchrun at N ranks, with the
# NCCL collective log (the 32-byte counter all-reduce of the BLER loop), and the BLER sweep module itself under torchrun.
N=${1:-2}
mkdir -p gpurun_out/r02
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL NCCL_DEBUG_FILE=gpurun_out/r02/nccl_n${N}.%h.%p.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 \
   2> gpurun_out/r02/bench_n${N}.err | tail -1 > gpurun_out/r02/bench_n${N}.json
tail -3 gpurun_out/r02/bench_n${N}.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02/bench_n${N}.json").read())
print("N",d["n_gpus"],"value",d["value"],"e2e",d["e2e"]["value"],"f64pageable",d["e2e"]["f64_pageable"]["value"])
print("config4",json.dumps(d["config4"]["points"]))
print("bler_loop",json.dumps(d["bler_loop"]))
PY
unset NCCL_DEBUG_FILE
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 -m ldpc_3gpp_matlab_b200.bler \
   --A 8424 --R 0.3333333333 --BG 1 --EsN0-start -0.6 --EsN0-delta 0.1 --target-block-errors 200 --target-BLER 1e-2 --batch 4096 \
   --out-dir gpurun_out/r02/results_n${N} > gpurun_out/r02/bler_sweep_n${N}.log 2>&1
grep -v NCCL gpurun_out/r02/bler_sweep_n${N}.log | tail -8
cat gpurun_out/r02/nccl_n${N}.*.log | grep -i "allreduce" | awk '{for(i=1;i<=NF;i++) if($i=="count") print $(i+1), $(i+2), $(i+3)}' | sort | uniq -c | sort -rn | head -8 > gpurun_out/r02/nccl_n${N}_allreduce_counts.txt
cat gpurun_out/r02/nccl_n${N}_allreduce_counts.txt
cat gpurun_out/r02/nccl_n${N}.*.log | grep -i "allreduce" | head -4
rm -f gpurun_out/r02/nccl_n${N}.*.log
