#!/bin/bash
# GPU session: bring-up of the CTA-pair engine mode, then A/B bench, then the full suite.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
timeout 120 python tools/pair_diag.py > gpurun_out/c1_diag.log 2>&1; echo "diag rc=$?" >> gpurun_out/c1_diag.log
timeout 600 python -m pytest tests/test_gpu_engines.py -q -x -k "not (102 or pair)" > gpurun_out/c1_eng_mcast.log 2>&1; echo "rc=$?" >> gpurun_out/c1_eng_mcast.log
timeout 600 python -m pytest tests/test_gpu_engines.py -q -k "102 or pair" > gpurun_out/c1_eng_pair.log 2>&1; echo "rc=$?" >> gpurun_out/c1_eng_pair.log
AEWN_ENGINE_MODE=mcast timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_mcast.json 2> gpurun_out/c1_bench_mcast.err
AEWN_ENGINE_MODE=pair timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/c1_bench_pair.json 2> gpurun_out/c1_bench_pair.err
AEWN_ENGINE_MODE=pair timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_engines.py > gpurun_out/c1_suite_pair.log 2>&1; echo "rc=$?" >> gpurun_out/c1_suite_pair.log
tail -3 gpurun_out/c1_diag.log gpurun_out/c1_eng_mcast.log gpurun_out/c1_eng_pair.log gpurun_out/c1_suite_pair.log
cat gpurun_out/c1_bench_mcast.json gpurun_out/c1_bench_pair.json | cut -c1-400
