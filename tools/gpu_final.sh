#!/bin/bash
# Round-end check on one B200: the whole GPU suite, smoke(), the default bench line (with the CPU baseline leg), the cfg3
# bench line and the reference arm.
mkdir -p gpurun_out; O=gpurun_out; T=${1:-fin}
timeout 900 python -m pytest tests -m gpu -q -x --durations=5 > $O/${T}_suite.log 2>&1; echo "rc=$?" >> $O/${T}_suite.log
tail -n 12 $O/${T}_suite.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; echo "rc=$?" >> $O/${T}_smoke.log; tail -n 3 $O/${T}_smoke.log
timeout 600 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 3 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
T=$T python - <<'PY'
import json,os
T=os.environ["T"]
for n in ("bench","bench_cfg3","bench_ref"):
    try:
        x=json.loads(open(f"gpurun_out/{T}_{n}.json").read().strip().splitlines()[-1])
        print(n, round(x["ms_per_step"],2), round(x["value"]), x.get("roofline",{}).get("frac"), x.get("clocks"), "e2e", x.get("e2e"), x.get("gpu_launches"), x.get("cpu_baseline"), x["config"].get("step_execution"))
        if "kernel_share" in x: print("   ", {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>0.5})
    except Exception as e:
        print(n, "ERR", e, open(f"gpurun_out/{T}_{n}.err").read()[-800:])
PY
