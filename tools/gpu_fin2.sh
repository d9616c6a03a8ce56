#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 400 python bench.py --no-cpu-baseline > $O/fin2_bench.json 2> $O/fin2_bench.err
timeout 300 python bench.py --workload cfg3 --steps 3 --warmup 3 > $O/fin2_cfg3.json 2> $O/fin2_cfg3.err
timeout 300 python bench_generate.py > $O/fin2_gen.json 2> $O/fin2_gen.err
python - <<'PY'
import json
for n in ("fin2_bench","fin2_cfg3","fin2_gen"):
    try:
        x=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, x.get("ms_per_step"), x.get("value"), x.get("config",{}).get("step_execution"), x.get("clocks"))
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/{n}.err").read()[-600:])
PY
