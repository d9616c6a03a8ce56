#!/bin/bash
# Round-3 evidence for profiles/: launch lists of one timed step (cfg2, cfg3), full ncu captures of the fused layer kernel
# (layer 0: dilation 1, layer 4: dilation 16), VQ latency, phase clock.
O=gpurun_out; T=${1:-r3}
mkdir -p $O
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${T}_launches_raw.csv $B > $O/${T}_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${T}_cfg3_launches_raw.csv $B --workload cfg3 --eager-cfg3 > $O/${T}_cfg3_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --nvtx --nvtx-include "aewn_timed" -k regex:grcc_fwd -s 0 -c 5 \
   -f -o $O/${T}_fused $B > $O/${T}_fused.log 2>&1
python profiles/vq_latency.py > $O/${T}_vq_latency.txt 2>&1
python profiles/gf_phase_clock.py 16 > $O/${T}_gf_phase_clock.txt 2>&1
ls -la $O | grep ${T}_
cat $O/${T}_vq_latency.txt
