#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $O/g1_graph.json 2> $O/g1_graph.err
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-graph > $O/g1_eager.json 2> $O/g1_eager.err
python - <<'PY'
import json
for n in ("graph","eager"):
    try:
        x=json.load(open(f"gpurun_out/g1_{n}.json"))
        print(n, round(x["ms_per_step"],2), round(x["value"]), x["config"]["step_execution"], x["config"]["final_loss"], "e2e", round(x["e2e"]["ms_per_step"],2), x["gpu_launches"], x["clocks"])
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/g1_{n}.err").read()[-1500:])
PY
