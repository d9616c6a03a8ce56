#!/bin/bash
# ncu evidence for the fused layer kernel: full captures (with source) of the first launches of a timed step.
# usage: tools/gpu_prof_fused.sh <tag> [skip] [count]
TAG=${1:-pf}; SKIP=${2:-0}; CNT=${3:-2}; O=gpurun_out
mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph"
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "aewn_timed" -k regex:grcc_fwd -s $SKIP -c $CNT \
   -f -o $O/${TAG}_fused $B > $O/${TAG}_fused.log 2>&1
tail -3 $O/${TAG}_fused.log
ls -la $O | grep $TAG
