#!/bin/bash
# GPU session 2: CTA-pair engine bring-up with early bail-out, A/B bench, cfg3 tests, full suite.
mkdir -p gpurun_out
O=gpurun_out
timeout 120 python tools/pair_diag.py > $O/c2_diag.log 2>&1; echo "diag rc=$?" >> $O/c2_diag.log
PAIR_OK=0
if [ "$(grep -c 'err_word=0 max_err=0.0' $O/c2_diag.log)" = "10" ]; then PAIR_OK=1; fi
echo "PAIR_OK=$PAIR_OK" >> $O/c2_diag.log
MODE=mcast
if [ $PAIR_OK = 1 ]; then
  timeout 240 python -m pytest tests/test_gpu_engines.py -q -x -k "102 or pair" > $O/c2_eng_pair.log 2>&1; rc=$?; echo "rc=$rc" >> $O/c2_eng_pair.log
  if [ $rc = 0 ]; then
    MODE=pair
    AEWN_ENGINE_MODE=pair timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > $O/c2_bench_pair.json 2> $O/c2_bench_pair.err
  fi
fi
echo "MODE=$MODE" >> $O/c2_diag.log
AEWN_ENGINE_MODE=$MODE timeout 400 python -m pytest tests/test_gpu_autoencoder.py -q -x -s > $O/c2_ae.log 2>&1; echo "rc=$?" >> $O/c2_ae.log
AEWN_ENGINE_MODE=$MODE timeout 300 python bench.py --workload cfg3 --steps 4 --warmup 3 > $O/c2_bench_cfg3.json 2> $O/c2_bench_cfg3.err
AEWN_ENGINE_MODE=$MODE timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_autoencoder.py --durations=8 > $O/c2_suite.log 2>&1; echo "rc=$?" >> $O/c2_suite.log
tail -4 $O/c2_diag.log $O/c2_eng_pair.log $O/c2_ae.log $O/c2_suite.log
cut -c1-300 $O/c2_bench_pair.json $O/c2_bench_cfg3.json
