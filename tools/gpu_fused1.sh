#!/bin/bash
# GPU session: bring-up of the fused layer kernel (aewn_grcc_fwd): unit tests first, then a short bench A/B.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-f1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/${T}_smi.log 2>&1
timeout 600 python -m pytest tests/test_gpu_fused_layer.py tests/test_gpu_vq_encoder.py -q > $O/${T}_fused.log 2>&1; rc=$?; echo "rc=$rc" >> $O/${T}_fused.log
grep -E "^(FAILED|ERROR|E  )|passed|failed" $O/${T}_fused.log | head -40
for F in 1 0; do
  AEWN_FUSED_FWD=$F timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${T}_bench_fused$F.json 2> $O/${T}_bench_fused$F.err
  tail -c 600 $O/${T}_bench_fused$F.err
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
