"""Where do the K1 value errors come from at large D?  Hypothesis: the tensor core's fp32 accumulation truncates, so the
3 D / 16 (fp32 form) or 2 D / 16 (fast form) accumulator updates leave a bias towards zero proportional to their count.
Prints, per (D, form), the max error and the mean SIGNED error along sign(f) on the largest values, relative to max|f|.
    python tools/acc_bias_probe.py          (B200)
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rff_oracle  # noqa: E402

pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
import bench  # noqa: E402

for D in (512, 2048, 8192):
    for kind in ("regression", "gaussian"):
        rng = np.random.default_rng(D)
        d, S, N = 6, 128, 8192
        W, b, Theta = bench.make_basis_arrays(D, d, S, 0.2, 0.0, 1.0, seed=D)
        if kind == "gaussian":
            Theta = rng.standard_normal((D, S))
        X = rng.uniform(0, 1, size=(N, d)).astype(np.float32)
        Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
        scale = np.abs(Fref).max(axis=0)
        big = np.abs(Fref) > 0.5 * scale
        for form in ("fp32", "fast"):
            basis = pkg.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision=form)
            F = basis.eval(X).cpu().numpy().T.astype(np.float64)
            err = (F - Fref) / scale
            signed = (err * np.sign(Fref))[big].mean()
            print(f"D={D:5d} {kind:10s} {form:4s} max|err|={np.abs(err).max():.2e} mean signed err on |f|>max/2: {signed:+.2e} "
                  f"(accumulator updates: {(3 if form == 'fp32' else 2) * D // 16})", flush=True)
