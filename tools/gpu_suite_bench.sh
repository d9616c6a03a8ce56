#!/bin/bash
# GPU session: full GPU suite, smoke(), then the default bench line (with baselines) and the cfg3 line.
mkdir -p gpurun_out
O=gpurun_out; T=${1:-s1}
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_suite.log 2>&1; echo "rc=$?" >> $O/${T}_suite.log
grep -E "^(FAILED|ERROR|E  )|passed|failed|rc=" $O/${T}_suite.log | cut -c1-400 | head -30
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 8 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err; tail -c 300 $O/${T}_bench.err
timeout 600 python bench.py --workload cfg3 --steps 6 --warmup 3 > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err; tail -c 300 $O/${T}_bench_cfg3.err
T=$T python - <<'PY'
import json,os
t=os.environ["T"]
for f in ("bench","bench_cfg3"):
    try:
        x=json.load(open(f"gpurun_out/{t}_{f}.json"))
        print(f, round(x["ms_per_step"],2), round(x["value"]), x["config"]["step_execution"][:40], "e2e", round(x["e2e"]["ms_per_step"],2), x["clocks"])
        r=x.get("roofline")
        if r: print("  frac", r["frac"], "ms/layer", r["ms_per_layer_fwd"], "inference", r["frac_inference"], r["ms_per_layer_inference"])
        ks=x.get("kernel_share")
        if ks: print("  outside engines", ks["outside_engines_frac"], {k:round(v,2) for k,v in ks["per_class_ms"].items() if v>0.3})
        if x.get("gpu_library_baseline"): print("  lib", {k:(v.get("ms_per_step") if isinstance(v,dict) else v) for k,v in x["gpu_library_baseline"].items() if k!="what"})
        if x.get("cpu_baseline"): print("  cpu", x["cpu_baseline"]["value"], x["cpu_baseline"]["cores"])
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err; tail -c 300 $O/${T}_bench_ref.err; python -c "
import json; x=json.load(open('$O/${T}_bench_ref.json')); print('reference arm:', x['value'], x['ms_per_step'], x['config'], x['cpu_baseline']['cores'])"
