#!/bin/bash
# bench line at N = 2 and N = 4 on one 4-GPU box (completes the 1 / 2 / 4 / 8 table of the committed sources)
O=gpurun_out/r02_scale24; mkdir -p $O
for N in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 50 --warmup 3 \
     2> $O/bench_n$N.err | tail -1 > $O/bench_n$N.json
  python - <<PY
import json
d=json.loads(open("$O/bench_n$N.json").read()); e=d["e2e"]
print("N",d["n_gpus"],"value",round(d["value"],2),"f16x2",round(d["f16x2"]["value"],2),"e2e",round(e["value"],2),round(e["f16_transport"]["value"],2),round(e["i8_transport"]["value"],2),
      "cfg4",[round(p["value"],1) for p in d["config4"]["points"]],[round(p["rank_imbalance"],4) for p in d["config4"]["points"]],"bler",round(d["bler_loop"]["frames_per_s"]))
PY
done
