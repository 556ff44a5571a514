#!/usr/bin/env python
"""Small pass over every kernel family for compute-sanitizer (memcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
ff, nat = pkg.fourier_features, pkg._native


def main():
    rng = np.random.default_rng(0)
    for (N, D, d, S, precision) in [(700, 130, 2, 40, "fp32"), (1500, 200, 6, 256, "fast"), (900, 64, 12, 130, "auto"),
                                    (5000, 320, 4, 300, "fast")]:
        X = rng.uniform(0, 1, (N, d)).astype(np.float32)
        W, b, T = rng.standard_normal((D, d)) / 0.3, rng.uniform(0, 6.28, D), rng.standard_normal((D, S))
        basis = ff.SharedBasis(W, b, T, device="cuda:0", precision=precision)
        basis.argmax(X)
        basis.eval(X)
        basis.argmax_key(X, idx_offset=5)
        if d % 2 == 0:
            basis.duel_copeland(X[:150, : d // 2].copy())
        pm = pkg.distributed.PeerMerge(S, "cuda:0")
        pm.argmax(basis, X, 3)
        pm.argmax(basis, torch.empty((0, d), device="cuda:0"), 0)
        ff.argmax_host(W, b, T, X)
        torch.cuda.synchronize()
        print("ok", (N, D, d, S, precision), basis.config(N)["S_chunk"], flush=True)
    Xl = rng.uniform(0, 1, (200000, 2)).astype(np.float32)               # several streamed chunks
    W, b, T = rng.standard_normal((64, 2)), rng.uniform(0, 6.28, 64), rng.standard_normal((64, 8))
    ff.argmax_host(W, b, T, Xl)
    gen = torch.Generator(device="cuda:0").manual_seed(1)
    fchi = torch.randn((128, 100, 5), device="cuda:0", generator=gen)
    fstar = torch.randn((128, 7), device="cuda:0", generator=gen)
    for mode in (nat.MPES_RANK, nat.MPES_PES):
        pkg.acquisitions._common.mpes_from_device_samples(fchi, fstar, 3 if mode == nat.MPES_RANK else 1, mode)
    x0 = rng.uniform(0, 1, (3, 5, 2))
    ff.ascend(x0, rng.standard_normal((3, 50, 2)), rng.uniform(0, 6, (3, 50, 1)), rng.standard_normal((3, 50, 1)), 0.0, 1.0,
              num_steps=5, device="cuda:0")
    torch.cuda.synchronize()
    print("sanitize smoke done")


if __name__ == "__main__":
    main()
