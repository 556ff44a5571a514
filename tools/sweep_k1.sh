#!/bin/bash
# K1 parity + timing with the producer-group knob (TKBO_NPG = 2, 3, 4; default: chosen per shape); writes gpurun_out/sweep_k1.log
mkdir -p gpurun_out
: > gpurun_out/sweep_k1.log
for g in default 2 3 4; do
  echo "== NPG=$g" >> gpurun_out/sweep_k1.log
  if [ "$g" = default ]; then timeout 300 python tests/gpu_probe.py k1 >> gpurun_out/sweep_k1.log 2>&1
  else TKBO_NPG=$g timeout 300 python tests/gpu_probe.py k1 >> gpurun_out/sweep_k1.log 2>&1; fi
done
cat gpurun_out/sweep_k1.log
