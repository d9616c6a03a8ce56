#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python bench.py --workload cfg5 --steps 4 --warmup 3 > $O/cfg5_bench.json 2> $O/cfg5_bench.err
python - <<'PY'
import json
try:
    x=json.loads(open("gpurun_out/cfg5_bench.json").read().strip().splitlines()[-1])
    print(round(x["ms_per_step"],2), round(x["value"]), x["roofline"]["frac"], x["clocks"], x["config"]["step_execution"], x["e2e"])
    print({k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>0.5})
except Exception as e:
    print("ERR", e, open("gpurun_out/cfg5_bench.err").read()[-1500:])
PY
