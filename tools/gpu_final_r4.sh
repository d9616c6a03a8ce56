#!/bin/bash
# Final evidence of the round: GPU suite + smoke + bench arms (tools/gpu_suite_bench.sh), loader latency, launch lists of one
# timed step (cfg2, cfg3), phase clock of the final build.
O=gpurun_out; T=${1:-r4}
bash tools/gpu_suite_bench.sh $T
python profiles/loader_latency.py > $O/${T}_loader_latency.txt 2>&1; cat $O/${T}_loader_latency.txt
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-graph"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${T}_launches_raw.csv $B > $O/${T}_launches_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "aewn_timed" --csv \
   --log-file $O/${T}_cfg3_launches_raw.csv $B --workload cfg3 --eager-cfg3 > $O/${T}_cfg3_launches_bench.log 2>&1
python profiles/gf_phase_clock.py 16 > $O/${T}_gf_phase_clock.txt 2>&1; grep -E "layer|whole" $O/${T}_gf_phase_clock.txt
ls -la $O | grep ${T}_ | awk '{print $5, $9}'
