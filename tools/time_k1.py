"""Median CUDA-event time of the fused argmax kernel on one bench shape (one GPU, rotating candidate buffers) and its value
error against the fp64 oracle on 8192 candidates.   TKBO_LIB=<variant .so> python tools/time_k1.py c2 [fp32|fast] [rows]"""
import importlib
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import rff_oracle  # noqa: E402

pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
form = sys.argv[2] if len(sys.argv) > 2 else "fp32"
if name.startswith("custom:"):          # custom:N,d,D,S  (uniform candidates in [0, 1]^d, lengthscale 0.2)
    cn, cd, cD, cS = (int(v) for v in name[7:].split(","))
    bench.WORKLOADS[name] = dict(N=cn, d=cd, D=cD, S=cS, lo=0.0, hi=1.0, ls=0.2, desc=name)
w = bench.WORKLOADS[name]
rows = int(sys.argv[3]) if len(sys.argv) > 3 else min(w["N"], 1 << 21)
dev = torch.device("cuda", 0)
W, b, Theta = bench.make_basis_arrays(w["D"], w["d"], w["S"], w["ls"], w["lo"], w["hi"])
basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev, precision=form)
X0 = torch.from_numpy(bench.candidates_host(name, 0, rows)).to(dev)
n_rot = int(min(16, max(2, -(-2 * bench.L2_BYTES // (rows * w["d"] * 4)))))
bufs = [X0] + [(X0 + v * 1e-4).clamp_(w["lo"], w["hi"]).contiguous() for v in range(1, n_rot)]
for i in range(5):
    basis.argmax(bufs[i % n_rot])
torch.cuda.synchronize()
reps = 40 if rows * w["D"] * w["S"] < (1 << 44) else 12
evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
evs[0].record()
for i in range(reps):
    basis.argmax(bufs[i % n_rot])
    evs[i + 1].record()
torch.cuda.synchronize()
ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(reps)]
clk = ""
try:                                                     # SM clock while the same loop keeps the GPU busy for ~0.3 s
    import pynvml
    pynvml.nvmlInit()
    hnd = pynvml.nvmlDeviceGetHandleByIndex(0)
    seen = []
    for _ in range(6):
        for i in range(max(1, int(50.0 / max(statistics.median(ms), 1e-3)))):
            basis.argmax(bufs[i % n_rot])
        seen.append(pynvml.nvmlDeviceGetClockInfo(hnd, pynvml.NVML_CLOCK_SM))
        torch.cuda.synchronize()
    clk = f"  sm_mhz {statistics.median(seen)}"
except Exception as exc:                                 # noqa: BLE001
    clk = f"  (no clock sample: {exc})"
Xs = bench.candidates_host(name, 0, 8192)
F = basis.eval(torch.from_numpy(Xs).to(dev)).cpu().numpy().T.astype(np.float64)
Fref = rff_oracle.rff_values_shared(Xs.astype(np.float64), W, b, Theta)
err = float((np.abs(F - Fref) / np.abs(Fref).max(axis=0)).max())
print(f"{os.environ.get('TKBO_LIB', 'libtkbo.so').split('/')[-1]:18s} {name} {form} rows={rows}: median {statistics.median(ms):.4f} ms "
      f"min {min(ms):.4f}  max|dv|/max|f| {err:.2e}  cfg {basis.config(rows)['producer_groups']}pg{clk}", flush=True)
