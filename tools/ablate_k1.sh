#!/bin/bash
# K1 timing ablations (TKBO_DEBUG bits: 1 = hi*hi MMA only, 2 = producers skip the cos, 4 = no Theta TMA, 8 = main MMA
# warps never pair slabs); numbers under an ablation are NOT results, they only show what the time is sensitive to.
# Bit 2 exists in the trace build only (python top-k-ranking-bayesian-optimization_b200/build.py --trace), so run against it:
export TKBO_LIB=${TKBO_LIB:-$(dirname "$0")/../top-k-ranking-bayesian-optimization_b200/libtkbo_trace.so}
mkdir -p gpurun_out
: > gpurun_out/ablate_k1.log
for dbg in 0 1 2 3 4 8; do
  echo "== DEBUG=$dbg" >> gpurun_out/ablate_k1.log
  TKBO_DEBUG=$dbg timeout 300 python tests/gpu_probe.py k1time >> gpurun_out/ablate_k1.log 2>&1
done
cat gpurun_out/ablate_k1.log
