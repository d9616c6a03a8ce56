#!/bin/bash
# ncu evidence for profiles/: (1) launch list of ONE timed step, (2) full captures of the dominant kernels.
# usage: tools/gpu_profile.sh <tag> <engine_mode>
TAG=${1:-r2}; MODE=${2:-pair}; O=gpurun_out
mkdir -p $O
export AEWN_ENGINE_MODE=$MODE
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${TAG}_launches_raw.csv $B > $O/${TAG}_launches_bench.log 2>&1
# layer-0 forward (conv+gate, res+skip) = the first two tgemm launches of the timed step
timeout 420 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "aewn_timed" -k regex:tgemm -c 2 \
   -f -o $O/${TAG}_fwd $B > $O/${TAG}_fwd.log 2>&1
# backward of the last layer: gate-derivative + data-gradient = the tgemm launches after 40 stack + 2 post forward + 2
# post backward ones; wide-unit weight gradients of the last layer = the first two wgradw launches
timeout 420 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "aewn_timed" -k regex:tgemm -s 44 -c 2 \
   -f -o $O/${TAG}_bwd $B > $O/${TAG}_bwd.log 2>&1
timeout 420 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "aewn_timed" -k regex:wgradw -c 2 \
   -f -o $O/${TAG}_wgrad $B > $O/${TAG}_wgrad.log 2>&1
ls -la $O | grep $TAG
