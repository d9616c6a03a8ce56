#!/bin/bash
# usage: tools/gpu_ab.sh TAG "ENV1=.. ENV2=.." "ENV..." ...   -- one bench.py run per environment string
mkdir -p gpurun_out; O=gpurun_out; T=$1; shift
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 240 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${T}_$i.json 2> $O/${T}_$i.err
  echo "$E" > $O/${T}_$i.env
done
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_*.json" % os.environ["T"])):
    try:
        x=json.load(open(f))
        print(open(f[:-5]+".env").read().strip(), "|", round(x["ms_per_step"],2), round(x["value"]), round(x["roofline"]["frac"],3), x["clocks"]["sm_mhz"], {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>1})
    except Exception as e:
        print(f, "ERR", e, open(f[:-5]+".err").read()[-300:])
PY
