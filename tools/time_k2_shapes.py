"""K2 time and achieved HBM bandwidth for other (C, k, mode) than the bench's c5: python tools/time_k2_shapes.py"""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
for (C, k, mode, Q, Sm, M) in [(2, 1, pkg._native.MPES_PES, 250000, 1000, 20), (2, 1, pkg._native.MPES_RANK, 250000, 1024, 20),
                               (3, 2, pkg._native.MPES_RANK, 160000, 1024, 20), (4, 2, pkg._native.MPES_RANK, 125000, 1024, 20),
                               (5, 3, pkg._native.MPES_RANK, 100000, 1024, 20)]:
    fstar = torch.randn((Sm, M), device=dev, generator=gen)
    fstar[:, 0] += 2.0
    fchi = torch.randn((Sm, Q, C), device=dev, generator=gen) * 0.5
    for _ in range(3):
        common.mpes_from_device_samples(fchi, fstar, k, mode)
    torch.cuda.synchronize()
    reps = 10
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        common.mpes_from_device_samples(fchi, fstar, k, mode)
        evs[r + 1].record()
    torch.cuda.synchronize()
    ms = statistics.median(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))
    gb = 4.0 * Sm * Q * C / 1e9
    print(f"C={C} k={k} mode={mode} Q={Q} S={Sm}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s ({gb:.2f} GB)", flush=True)
    del fchi
