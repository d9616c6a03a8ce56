#!/bin/bash
# GPU session: bring-up of the fused layer kernel (aewn_grcc_fwd): unit test first, then a short bench A/B.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-f1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${T}_smi.log 2>&1
timeout 420 python -m pytest tests/test_gpu_fused_layer.py -x -q > $O/${T}_fused.log 2>&1; rc=$?; echo "rc=$rc" >> $O/${T}_fused.log
tail -n 30 $O/${T}_fused.log
if [ $rc != 0 ]; then exit 1; fi
for F in 1 0; do
  AEWN_FUSED_FWD=$F timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${T}_bench_fused$F.json 2> $O/${T}_bench_fused$F.err
  tail -c 600 $O/${T}_bench_fused$F.err
done
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % os.environ["T"])):
    try:
        x=json.load(open(f))
        print(f, round(x["ms_per_step"],2), round(x["value"]), x["roofline"], x["clocks"]["sm_mhz"], {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>0.5})
    except Exception as e:
        print(f, "ERR", e)
PY
