#!/bin/bash
# timing experiments on the fused layer kernel: one phase-clock run per AEWN_GF_DBG mask
mkdir -p gpurun_out
for M in "$@"; do
  echo "=== AEWN_GF_DBG=$M"
  AEWN_GF_DBG=$M python profiles/gf_phase_clock.py 4 > gpurun_out/dbg_$M.txt 2>&1
  head -1 gpurun_out/dbg_$M.txt; grep "tile 2" gpurun_out/dbg_$M.txt
done
