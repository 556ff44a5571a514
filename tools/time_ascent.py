#!/usr/bin/env python
"""Time the on-device sign-gradient ascent (ff.py:139-157) at the c1 shape: count=20, n_init=50, D=1000, d=1."""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
rng = np.random.default_rng(1)
count, n_init, D, d = 20, 50, 1000, 1
W = rng.standard_normal((count, D, d)) / 0.2
b = rng.uniform(0, 2 * np.pi, (count, D, 1))
theta = rng.standard_normal((count, D, 1))
x0 = rng.uniform(0, 1, (count, n_init, d))
for steps in (300, 3000):
    pkg.fourier_features.ascend(x0, W, b, theta, 0.0, 1.0, num_steps=10, device="cuda:0")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, ran = pkg.fourier_features.ascend(x0, W, b, theta, 0.0, 1.0, num_steps=steps, device="cuda:0")
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"num_steps={steps}: ran {ran} steps in {1e3 * (t1 - t0):.2f} ms ({1e6 * (t1 - t0) / max(ran, 1):.1f} us/step)")
