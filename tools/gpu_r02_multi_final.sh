#!/bin/bash
# r02 final multi-GPU visit: the default bench line under torchrun at N ranks (side keys config4 / bler_loop), the BLER sweep
# module under torchrun with NCCL's log of the counter all-reduce, and BASELINE config 5 (the sweep table) on N GPUs.
N=${1:-8}
O=gpurun_out/r02_n${N}_final; mkdir -p $O
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=COLL NCCL_DEBUG_FILE=$O/nccl.%h.%p.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 50 --warmup 3 \
   2> $O/bench.err | tail -1 > $O/bench_n${N}.json
tail -2 $O/bench.err
python - <<PY
import json
d=json.loads(open("$O/bench_n${N}.json").read())
print("N",d["n_gpus"],"value",d["value"],"f16x2",d["f16x2"]["value"],"e2e",d["e2e"]["value"],"f64pageable",d["e2e"]["f64_pageable"]["value"],"clocks",d["clocks"])
print("config4",json.dumps(d["config4"]["points"]))
print("bler_loop",json.dumps(d["bler_loop"]))
PY
unset NCCL_DEBUG_FILE
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 -m ldpc_3gpp_matlab_b200.bler \
   --A 8424 --R 0.3333333333 --BG 1 --EsN0-start -0.6 --EsN0-delta 0.1 --target-block-errors 200 --target-BLER 1e-2 --batch 4096 \
   --out-dir $O/results > $O/bler_sweep.log 2>&1
grep -v NCCL $O/bler_sweep.log | tail -6
cat $O/nccl.*.log | grep -i "allreduce" | awk '{for(i=1;i<=NF;i++) if($i=="count") print $(i+1), $(i+2), $(i+3)}' | sort | uniq -c | sort -rn | head -8 > $O/nccl_allreduce_counts.txt
cat $O/nccl_allreduce_counts.txt
rm -f $O/nccl.*.log
unset NCCL_DEBUG NCCL_DEBUG_SUBSYS
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/sweep.py --mb 400 --out $O/sweep_${N}gpu > $O/sweep.log 2>&1
grep -E "^\| (2|52|208|384) " $O/sweep_${N}gpu.md
