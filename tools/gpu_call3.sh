#!/bin/bash
# GPU session: warp-convergent issue loops -- correctness (diag + engine tests), A/B/auto bench, full suite.
mkdir -p gpurun_out
O=gpurun_out
T=${1:-c3}
timeout 240 python -u tools/pair_diag.py > $O/${T}_diag.log 2>&1; echo "diag rc=$?" >> $O/${T}_diag.log
if [ "$(grep -c 'err_word=0 max_err=0.0' $O/${T}_diag.log)" != "14" ]; then echo "DIAG FAILED"; cat $O/${T}_diag.log; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_engines.py -q -x > $O/${T}_eng.log 2>&1; rc=$?; echo "rc=$rc" >> $O/${T}_eng.log
tail -3 $O/${T}_eng.log
if [ $rc != 0 ]; then exit 1; fi
for M in ${MODES:-narrow auto}; do
  if [ $M = narrow ]; then export AEWN_ENGINE_MODE=auto AEWN_WIDE_WGRAD=0; else export AEWN_ENGINE_MODE=$M AEWN_WIDE_WGRAD=1; fi
  timeout 200 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/${T}_bench_$M.json 2> $O/${T}_bench_$M.err
done
unset AEWN_ENGINE_MODE AEWN_WIDE_WGRAD
timeout 500 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_engines.py > $O/${T}_suite.log 2>&1; echo "rc=$?" >> $O/${T}_suite.log
tail -n 3 $O/${T}_suite.log
T=$T python - <<'PY'
import json,glob,os
for f in sorted(glob.glob("gpurun_out/%s_bench_*.json" % os.environ["T"])):
    try:
        x=json.load(open(f))
        print(f, round(x["ms_per_step"],2), round(x["value"]), round(x["roofline"]["frac"],3), x["clocks"]["sm_mhz"], {k:round(v,2) for k,v in x["kernel_share"]["per_class_ms"].items() if v>1})
    except Exception as e:
        print(f, "ERR", e)
PY
