#!/bin/bash
# GPU session: fused layer kernel unit tests, then bench A/B over environment settings (one bench run per argument).
# usage: tools/gpu_fused2.sh TAG "ENV=.. ENV=.." "ENV=.." ...
mkdir -p gpurun_out
O=gpurun_out
T=$1; shift
timeout 600 python -m pytest tests/test_gpu_fused_layer.py -q > $O/${T}_fused.log 2>&1; rc=$?; echo "rc=$rc" >> $O/${T}_fused.log
grep -E "^(FAILED|ERROR|E  )|passed|failed" $O/${T}_fused.log | cut -c1-300 | head -20
if [ $rc != 0 ]; then exit 1; fi
i=0
for E in "$@"; do
  env $E timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${T}_bench_$i.json 2> $O/${T}_bench_$i.err
  echo "[$i] $E"; tail -c 300 $O/${T}_bench_$i.err
  i=$((i+1))
done
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % os.environ["T"])):
    try:
        x=json.load(open(f))
        r=x["roofline"]
        print(f, round(x["ms_per_step"],2), round(x["value"]), round(r["frac"],3), round(r["ms_per_layer_fwd"],4), x["clocks"]["sm_mhz"], {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>0.5})
        print("   per-layer GB/s", r["per_layer_gbs"])
    except Exception as e:
        print(f, "ERR", e)
PY
