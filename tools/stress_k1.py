#!/usr/bin/env python
"""Random-shape stress of the fused RFF kernel against the fp64 oracle (argmax, store-F and duel modes).
    python tools/stress_k1.py [n_cases] [seed]
Run under `timeout`: a pipeline deadlock would show up as a hang."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import mpes_oracle, rff_oracle  # noqa: E402

pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    worst = 0.0
    for case in range(n_cases):
        d = int(rng.choice([1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16]))
        D = int(rng.choice([1, 17, 64, 65, 128, 200, 320, 700]))
        S = int(rng.choice([1, 5, 32, 33, 64, 100, 128, 129, 256, 300, 520]))
        N = int(rng.choice([1, 31, 128, 129, 1000, 5000, 19000, 40000]))
        if N * D * S > 6e9:
            N = max(1, int(6e9 / (D * S)))
        X = rng.uniform(-1, 1, size=(N, d)).astype(np.float32)
        W = rng.standard_normal((D, d)) / rng.uniform(0.3, 1.5)
        b = rng.uniform(0, 2 * np.pi, D)
        Theta = rng.standard_normal((D, S)) * rng.uniform(0.01, 50.0, size=(1, S))
        basis = pkg.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
        cfg = basis.config(N)
        Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
        # normalise by the larger of max|f| and the weight norm (a single candidate can sit at a zero of f)
        scale = np.maximum(np.abs(Fref).max(axis=0), np.linalg.norm(Theta * np.sqrt(2.0 / D), axis=0) + 1e-300)
        F = basis.eval(X).cpu().numpy().T.astype(np.float64)
        err = float(np.max(np.abs(F - Fref) / scale))
        val, idx = basis.argmax(X)
        idx = idx.cpu().numpy()
        got = Fref[idx, np.arange(S)]
        gap = float(np.max((Fref.max(axis=0) - got) / scale))
        verr = float(np.max(np.abs(val.cpu().numpy() - Fref.max(axis=0)) / scale))
        msg = f"case {case:3d} N={N:6d} D={D:4d} d={d:2d} S={S:4d} SC={cfg['S_chunk']:3d} ring={cfg['tmem_ring']} npg={cfg['producer_groups']} pair={cfg['cta_pairs']} " \
              f"stages={cfg['stages']:2d}: eval {err:.1e} argmax-gap {gap:.1e} val {verr:.1e}"
        if d % 2 == 0 and N <= 5000:
            n = min(N, 150)
            Xb = X[:n, : d // 2].copy()
            score, win = basis.duel_copeland(Xb)
            combs = pkg.acquisitions.dts.combinations(Xb.astype(np.float64))
            Fd = rff_oracle.rff_values_shared(combs, W, b, Theta)
            ref = np.stack([mpes_oracle.soft_copeland_scores(Fd[:, s]) for s in range(S)])
            derr = float(np.max(np.abs(score.cpu().numpy() - ref)))
            msg += f" duel {derr:.1e}"
            worst = max(worst, derr * 10)
        print(msg, flush=True)
        worst = max(worst, err, gap, verr)
        assert err < 1e-4 and gap < 1e-4 and verr < 1e-4, msg
        del basis
    print(f"stress ok: {n_cases} cases, worst normalised error {worst:.2e}")


if __name__ == "__main__":
    main()
