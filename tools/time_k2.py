"""K2 (mpes_all_perms_kernel<5,3>) time against the number of query sets: shows the cost of the last, partly filled round
of 32-query tiles over the resident blocks (148 SMs x 4 blocks).   python tools/time_k2.py [Q ...]"""
import importlib
import os
import statistics
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
dev = torch.device("cuda", 0)
Sm, C, k, M = 1024, 5, 3, 20
gen = torch.Generator(device=dev).manual_seed(0)
fstar = torch.randn((Sm, M), device=dev, generator=gen)
fstar[:, 0] += 2.0                       # skewed groups, like the bench input
for Q in [int(a) for a in sys.argv[1:]] or [94720, 100000, 104192, 113664]:
    fchi = torch.randn((Sm, Q, C), device=dev, generator=gen) * 0.5
    for _ in range(3):
        common.mpes_from_device_samples(fchi, fstar, k, pkg._native.MPES_RANK)
    torch.cuda.synchronize()
    reps = 15
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        common.mpes_from_device_samples(fchi, fstar, k, pkg._native.MPES_RANK)
        evs[r + 1].record()
    torch.cuda.synchronize()
    ms = statistics.median(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))
    tiles = (Q + 31) // 32
    print(f"Q={Q:7d} tiles={tiles:5d} rounds={tiles / 592:.2f}: {ms:.4f} ms  ({1e6 * ms / tiles:.1f} ns per tile)", flush=True)
    del fchi
