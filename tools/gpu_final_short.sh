#!/bin/bash
# Short final rehearsal: GPU suite, smoke, default bench, cfg3 bench, launch list of one cfg2 step.
O=gpurun_out; T=${1:-r7}
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q > $O/${T}_suite.log 2>&1; echo "rc=$?" >> $O/${T}_suite.log
grep -E "^(FAILED|ERROR|E  )|passed|failed|rc=" $O/${T}_suite.log | cut -c1-300 | head -20
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 8 --warmup 3 > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --workload cfg3 --steps 6 --warmup 3 --no-cpu-baseline --no-gpu-baseline > $O/${T}_bench_cfg3.json 2> $O/${T}_bench_cfg3.err
T=$T python - <<'PY'
import json,os
t=os.environ["T"]
for f in ("bench","bench_cfg3"):
    try:
        x=json.load(open(f"gpurun_out/{t}_{f}.json"))
        print(f, round(x["ms_per_step"],2), round(x["value"]), "e2e", round(x["e2e"]["ms_per_step"],2), x["clocks"], round(x["roofline"]["frac"],3), x["roofline"]["ms_per_layer_fwd"], x["roofline"].get("frac_inference"))
        print("  ", {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>0.3}, x.get("gpu_library_baseline",{}).get("speedup_vs_tf32"))
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${T}_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-graph > $O/${T}_launches_bench.log 2>&1
ls $O | grep ${T}_ | tr '\n' ' '
