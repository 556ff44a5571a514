#!/bin/bash
# K1 timing sweep over the experiment knobs (TKBO_TILES, TKBO_NPG); writes gpurun_out/sweep_k1.log
mkdir -p gpurun_out
: > gpurun_out/sweep_k1.log
for t in 1; do for g in 2 3 4; do
  echo "== NPG=$g" >> gpurun_out/sweep_k1.log
  TKBO_NPG=$g timeout 300 python tests/gpu_probe.py k1 >> gpurun_out/sweep_k1.log 2>&1
done; done
cat gpurun_out/sweep_k1.log
