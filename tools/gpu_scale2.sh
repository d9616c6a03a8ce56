#!/bin/bash
# 2-GPU session: cfg2 weak scaling and cfg4 (= cfg3 per GPU, grads + EMA statistics in one all-reduce)
mkdir -p gpurun_out; O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
timeout 400 $TR bench.py --gpus 2 --steps 6 --warmup 3 > $O/s2_cfg2.json 2> $O/s2_cfg2.err
timeout 400 $TR bench.py --gpus 2 --steps 4 --warmup 3 --workload cfg3 > $O/s2_cfg4.json 2> $O/s2_cfg4.err
timeout 200 python bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/s2_ref.json 2> $O/s2_ref.err
python - <<'PY'
import json
for n in ("s2_cfg2","s2_cfg4","s2_ref"):
    try:
        x=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, x.get("n_gpus"), round(x["ms_per_step"],2), round(x["value"]), x["config"].get("step_execution"), x.get("clocks"), x.get("e2e"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-1200:])
PY
