#!/bin/bash
# A/B of environment settings with the single-layer phase-clock script (prints the layer time only), then the unit tests
mkdir -p gpurun_out
for E in "$@"; do
  for D in 4 1 512; do
    echo "[$E] dil=$D: $(env $E python profiles/gf_phase_clock.py $D 2>&1 | head -1)"
  done
done
timeout 300 python -m pytest tests/test_gpu_fused_layer.py -q 2>&1 | tail -2
