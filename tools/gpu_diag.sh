#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
for part in wgradw tgemm wgrad; do
  timeout 60 python -u tools/pair_diag.py $part > $O/d_$part.log 2>&1; echo "rc=$?" >> $O/d_$part.log
done
tail -n 12 $O/d_wgradw.log $O/d_tgemm.log $O/d_wgrad.log
