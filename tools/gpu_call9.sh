#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; T=c9
timeout 240 python -u tools/pair_diag.py > $O/${T}_diag.log 2>&1
if [ "$(grep -c 'err_word=0 max_err=0.0' $O/${T}_diag.log)" != "14" ]; then echo "DIAG FAILED"; tail -20 $O/${T}_diag.log; exit 1; fi
timeout 300 python -m pytest tests/test_gpu_engines.py tests/test_gpu_decoder.py tests/test_gpu_fullsize.py -q -x > $O/${T}_tests.log 2>&1; rc=$?; tail -3 $O/${T}_tests.log
if [ $rc != 0 ]; then tail -40 $O/${T}_tests.log; exit 1; fi
bash tools/gpu_ab.sh ${T}ab "AEWN_MERGE_DGRAD=0" "AEWN_MERGE_DGRAD=1" "AEWN_MERGE_DGRAD=0" "AEWN_MERGE_DGRAD=1"
