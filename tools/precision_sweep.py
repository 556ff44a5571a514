#!/usr/bin/env python
"""Time the fused argmax kernel with the fp32 and the fast (e4m3 correction) contraction on a few shapes:
    python tools/precision_sweep.py
Shows where the fast form pays (sample chunks >= 128, the tensor-bound regime: SharedBasis(precision="auto", f_scale=...)
only considers it for S > 96)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")


def main():
    rng = np.random.default_rng(0)
    for (N, D, d, S) in [(1 << 19, 1024, 2, 64), (1 << 19, 1024, 2, 96), (1 << 19, 1024, 2, 128), (1 << 19, 1024, 6, 128),
                         (1 << 19, 1024, 2, 160), (1 << 19, 1024, 6, 192), (1 << 18, 2048, 4, 256), (1 << 18, 2048, 12, 256)]:
        X = torch.from_numpy(rng.uniform(0, 1, (N, d)).astype(np.float32)).cuda()
        W, b, T = rng.standard_normal((D, d)) / 0.3, rng.uniform(0, 6.28, D), rng.standard_normal((D, S))
        out = []
        for precision in ("fp32", "fast"):
            basis = pkg.fourier_features.SharedBasis(W, b, T, device="cuda:0", precision=precision)
            cfg = basis.config(N)
            for _ in range(3):
                basis.argmax(X)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                basis.argmax(X)
            e1.record()
            torch.cuda.synchronize()
            out.append(e0.elapsed_time(e1) / 10)
        print(f"N={N} D={D} d={d} S={S} SC={cfg['S_chunk']} ring={cfg['tmem_ring']} npg={cfg['producer_groups']} "
              f"stages={cfg['stages']}+{cfg['w_stages']} share={cfg['producers_share_slab']}: fp32 {out[0]:.3f} ms, fast {out[1]:.3f} ms "
              f"({out[0] / out[1]:.2f}x)", flush=True)


if __name__ == "__main__":
    main()
