import importlib, os, sys, torch
sys.path.insert(0, "/root/repo")
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
dev = torch.device("cuda", 0)
gen = torch.Generator(device=dev).manual_seed(0)
C, k, Q, Sm, M = 2, 1, 250000, 1000, 20
fstar = torch.randn((Sm, M), device=dev, generator=gen); fstar[:, 0] += 2.0
fchi = torch.randn((Sm, Q, C), device=dev, generator=gen) * 0.5
for _ in range(2):
    common.mpes_from_device_samples(fchi, fstar, k, pkg._native.MPES_PES)
torch.cuda.synchronize()
